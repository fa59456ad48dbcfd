"""Aggregate an ncu SASS source page by CUDA source line, using nvdisasm line info of the local build.

usage: python tools/ncu_by_line.py <report.ncu-rep | source.csv> <kernel mangled-name substring> [top N]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "tsim_b200", "libtsim_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")), key=os.path.getsize)  # the host-only object has an (empty) cubin too
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
# find the text section of the kernel
sec = None
off2line = {}
cur = None
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        sec = m.group(1)
        continue
    if sec is None or kname not in sec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
if rep.endswith(".csv"):  # a source page exported on the GPU box (ncu -i rep --page source --csv)
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ia, ii, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
body = rows[hi + 1 :]
base = int(body[0][ia], 16)
by = collections.Counter()
samp = collections.Counter()
tot = 0
for r in body:
    off = int(r[ia], 16) - base
    n = int(r[ii] or 0)
    key = off2line.get(off, (None, "?"))[0]
    by[key] += n
    samp[key] += int(r[isamp] or 0)
    tot += n
print(f"total warp instructions {tot}")
src_cache = {}
for key, n in by.most_common(top):
    text = ""
    if key:
        for d in ("tsim_b200/csrc",):
            p = os.path.join(root, d, key[0])
            if os.path.exists(p):
                src_cache.setdefault(p, open(p).read().splitlines())
                text = src_cache[p][key[1] - 1].strip()[:100]
    print(f"{n/tot*100:5.1f}%  samp {samp[key]:6d}  {key}  {text}")
