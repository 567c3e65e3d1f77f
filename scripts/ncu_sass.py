"""SASS-level view of a source line range: python scripts/ncu_sass.py rep kernel unit lo hi"""
import collections, csv, io, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern, unit, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
work = os.path.join(ROOT, "scratch", "sass"); os.makedirs(work, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "lzfse_rust_b200", "liblzfse_b200.so")], cwd=work, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(work, unit + ".sm_100a.cubin")], capture_output=True, text=True).stdout
line_of, cur, inside = {}, None, False
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."): inside = kern in l
    if not inside: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: line_of[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[h0]
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None; addrs = sorted(line_of)
# address range covering the requested lines (first to last instruction attributed to them), printed contiguously
sel = [a for a in addrs if line_of[a] and line_of[a][0].startswith(unit) and lo <= line_of[a][1] <= hi]
amin, amax = min(sel), max(sel)
for r in rows[h0 + 1:]:
    if not r or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    off = a - base
    if off < amin or off > amax: continue
    st = sorted(((int(float(r[i] or 0)), h[6:]) for i, h in stall_cols), reverse=True)[:2]
    print("%05x L%-5s %6s samp %9s ex  %-60s %s" % (off, line_of.get(off, ("", 0))[1] if line_of.get(off) else "?", r[isamp], r[iex], r[1].strip()[:60], [s for s in st if s[0]]))
