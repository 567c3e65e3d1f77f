#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference checkout (run in the build container only).

/root/reference does not exist on the GPU box, so everything the tests need is committed here:
  data/{snappy,mutate,special}/*.lzfse + *.hash   -- the reference's own decode fixtures
        (frames produced by Apple's C lzfse encoder; .hash = SHA-256 of the C decoder's output,
        /root/reference/scripts/hash.sh:9-15; used by /root/reference/test/src/data.rs:84-97)
  data/synth/{random,word08,repl01,repl16,reps08,repsin}.lzfse -- a few of the bench-only frames
  lmd_sha256.json   -- SHA-256 of each data/snappy/lmdy_output/*.lmd text dump (the dumps are 3.5 MB)
  encoder_kat.json  -- byte-exact encoder KATs parsed out of
        /root/reference/src/encode/frontend_bytes.rs:455-531 and src/encode/mod.rs:50-54
"""
import glob
import hashlib
import json
import os
import re
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    for d in ("snappy", "mutate", "special"):
        os.makedirs(f"{HERE}/data/{d}", exist_ok=True)
        for f in glob.glob(f"{REF}/data/{d}/*.lzfse") + glob.glob(f"{REF}/data/{d}/*.hash"):
            shutil.copy(f, f"{HERE}/data/{d}/")
    os.makedirs(f"{HERE}/data/synth", exist_ok=True)
    for n in ("random", "word08", "repl01", "repl16", "reps08", "repsin"):
        shutil.copy(f"{REF}/data/synth/{n}.lzfse", f"{HERE}/data/synth/")
    lmd = {}
    for f in sorted(glob.glob(f"{REF}/data/snappy/lmdy_output/*.lmd")):
        lmd[os.path.basename(f)[:-4]] = hashlib.sha256(open(f, "rb").read()).hexdigest()
    json.dump(lmd, open(f"{HERE}/lmd_sha256.json", "w"), indent=1, sort_keys=True)
    src = open(f"{REF}/src/encode/frontend_bytes.rs").read()
    kat = {}
    for n in (0, 1, 20, 21, 4096, 4097):
        m = re.search(r"fn zero_%d\(\).*?check_output\(\s*&\[0; %d\],\s*&\[(.*?)\],?\s*\)" % (n, n), src, re.S)
        kat["zero_%d" % n] = {"input_zero_len": n, "frame_hex": "".join(re.findall(r"0x([0-9A-Fa-f]{2})", m.group(1))).lower()}
    doc = open(f"{REF}/src/encode/mod.rs").read()
    m = re.search(r"assert_eq!\(enc, &\[(.*?)\]\);", doc, re.S)
    kat["test"] = {"input_ascii": "test", "frame_hex": "".join(re.findall(r"0x([0-9A-Fa-f]{2})", m.group(1))).lower()}
    json.dump(kat, open(f"{HERE}/encoder_kat.json", "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
