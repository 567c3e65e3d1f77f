"""Joins an ncu SASS source page with nvdisasm line info: stall samples and executed instructions per CUDA source line.

  python scripts/ncu_lines.py gpurun_out/x.ncu-rep k_enc_find [encode] [top]

The cubin is taken from the in-tree liblzfse_b200.so (must be the build that was profiled)."""
import collections, csv, io, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
unit = sys.argv[3] if len(sys.argv) > 3 else "encode"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lo, hi_line = (int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (0, 1 << 30)  # optional line range: per-line stall reasons
work = os.path.join(ROOT, "scratch", "sass"); os.makedirs(work, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "lzfse_rust_b200", "liblzfse_b200.so")], cwd=work, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(work, unit + ".sm_100a.cubin")], capture_output=True, text=True).stdout
line_of, cur, inside = {}, None, False
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        inside = kern in l
    if not inside: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + os.environ.get("NCU_KERNEL_REGEX", kern)], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isamp, iex, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]; base = None
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
stalls = collections.defaultdict(collections.Counter)
for r in rows[hi + 1:]:
    if len(r) <= ithr or not r[ia]: continue
    try: a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except ValueError: continue
    if base is None: base = a
    key = line_of.get(a - base, (None, "?"))[0]
    v = [int(float(r[isamp] or 0)), int(float(r[iex] or 0)), int(float(r[ithr] or 0))]
    for k in range(3): agg[key][k] += v[k]; tot[k] += v[k]
    for i, h in stall_cols:
        if r[i] not in ("", "0"): stalls[key][h[6:]] += int(float(r[i]))
print("total samples %d, warp-instr %d, thread-instr %d (%.1f threads/instr)" % (tot[0], tot[1], tot[2], tot[2] / max(tot[1], 1)))
src_cache = {}
def text(key):
    if not key: return ""
    f, n = key
    p = os.path.join(ROOT, "lzfse_rust_b200", "csrc", f)
    if p not in src_cache: src_cache[p] = open(p).read().splitlines() if os.path.exists(p) else []
    return src_cache[p][n - 1].strip()[:110] if 0 < n <= len(src_cache[p]) else ""
items = [kv for kv in agg.items() if kv[0] and lo <= kv[0][1] <= hi_line] if len(sys.argv) > 6 else list(agg.items())
for key, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% samp %5.1f%% instr  thr/instr %4.1f  %s:%s  %s" % (100.0 * v[0] / max(tot[0], 1), 100.0 * v[1] / max(tot[1], 1), v[2] / max(v[1], 1), key[0] if key else "?", key[1] if key else "?", text(key)))
    if len(sys.argv) > 6: print("        ", dict(stalls[key].most_common(4)))
