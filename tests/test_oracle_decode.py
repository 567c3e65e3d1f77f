"""Pins the CPU oracle's decoder against every golden vector the reference's tests hold for
decode_bytes: fixture SHA-256s (test/src/data.rs:84-97), the LMD dumps
(data/snappy/lmdy_output), the negative vector, and the mutate family (test/src/mutate_0..7.rs:
a mutated frame must produce a status, never crash, and the pristine frame must still hash right)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
import testkit as tk


@pytest.mark.parametrize("name,frame,digest", tk.golden_frames(), ids=[g[0] for g in tk.golden_frames()])
def test_fixture_sha256(name, frame, digest):
    st, out = ob.decode(frame)
    if name == "special/null.vx2":
        # hostile header: 0-byte weight table (fse/weights.rs:98-100)
        assert st == 30 and out == b""
        return
    assert st == 0
    assert tk.sha256(out) == digest
    st, raw, _ = ob.probe(frame)
    assert st == 0 and raw == len(out)


def test_lmd_traces(golden_dir):
    want = json.load(open(os.path.join(golden_dir, "lmd_sha256.json")))
    assert len(want) == 12
    for name, digest in want.items():
        frame = open(os.path.join(golden_dir, "data", "snappy", name + ".lzfse"), "rb").read()
        st, out, tr = ob.decode_trace(frame, 1 << 20)
        assert st == 0
        lines = []
        for l, m, d in tr:
            if l != ob.NO_L:
                lines.append("L:%d" % l)
            if m:
                lines.append("M:%d" % m)
                lines.append("D:%d" % d)
        text = ("\n".join(lines) + "\n").encode()
        assert hashlib.sha256(text).hexdigest() == digest, name


def test_synth_frames_decode(golden_dir):
    for n in ("random", "word08", "repl01", "repl16", "reps08", "repsin"):
        frame = open(os.path.join(golden_dir, "data", "synth", n + ".lzfse"), "rb").read()
        st, out = ob.decode(frame)
        assert st == 0 and len(out) >= 65000


def test_eos_only_and_trailing_garbage():
    assert ob.decode(b"bvx$") == (0, b"")                      # decode/decoder.rs:79-98
    assert ob.decode(b"bvx$\0")[0] == 6                        # PayloadOverflow
    assert ob.decode(b"bvx")[0] == 7                           # PayloadUnderflow
    assert ob.decode(b"bvxq\0\0\0\0")[0] == 1                  # BadBlock
    assert ob.decode(b"")[0] == 7


def test_opcode_table():
    # vn/constants.rs:39-72 restated as a 256-entry string: one letter per class
    # S=SmlD E=Eos L=LrgD N=Nop U=Udef P=PreD M=MedD l=SmlL/LrgL m=SmlM/LrgM
    rows = ["SSSSSSEL", "SSSSSSNL", "SSSSSSNL", "SSSSSSUL", "SSSSSSUL", "SSSSSSUL", "SSSSSSUL", "SSSSSSUL"]
    rows += ["SSSSSSPL"] * 6 + ["UUUUUUUU"] * 2 + ["SSSSSSPL"] * 4 + ["MMMMMMMM"] * 4 + ["SSSSSSPL"] * 2 + ["UUUUUUUU"] * 2
    rows += ["llllllll"] * 2 + ["mmmmmmmm"] * 2
    tbl = "".join(rows)
    assert len(tbl) == 256
    cls = {0: "l", 1: "l", 2: "m", 3: "m", 4: "P", 5: "S", 6: "M", 7: "L", 8: "E", 9: "U", 10: "N"}
    for b in range(256):
        c = ob.lib().orc_vn_op_class(b)
        assert cls[c] == tbl[b], hex(b)
    assert ob.lib().orc_vn_op_class(0xE0) == 1 and ob.lib().orc_vn_op_class(0xF0) == 3


def _mutations(frame, seed, n):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        b = bytearray(frame)
        kind = rng.integers(0, 5)
        if kind == 0:
            i = rng.integers(0, len(b)); b[i] ^= 1 << rng.integers(0, 8)
        elif kind == 1:
            i = rng.integers(0, len(b)); b[i] = rng.integers(0, 256)
        elif kind == 2:
            i = rng.integers(0, max(1, len(b) - 4)); b[i:i + 4] = rng.integers(0, 256, 4, dtype=np.uint8).tobytes()
        elif kind == 3:
            b = b[: rng.integers(0, len(b))]
        else:
            b += rng.integers(0, 256, rng.integers(1, 9), dtype=np.uint8).tobytes()
        yield bytes(b)


@pytest.mark.parametrize("name", ["raw", "vx1", "vx2", "vxn"])
def test_mutate_never_crashes(golden_dir, name):
    frame = open(os.path.join(golden_dir, "data", "mutate", name + ".lzfse"), "rb").read()
    digest = open(os.path.join(golden_dir, "data", "mutate", name + ".hash"), "rb").read()
    n_err = 0
    for m in _mutations(frame, 1234, 400):
        st, out = ob.decode(m, cap=1 << 16)
        assert 0 <= st < 64
        n_err += st != 0
    assert n_err > 100
    st, out = ob.decode(frame)
    assert st == 0 and tk.sha256(out) == digest
