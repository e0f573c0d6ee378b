"""Attribute ncu warp-stall samples to CUDA source lines: joins `ncu --page source` (SASS level) with
nvdisasm line info of the built library.  usage: ncu_lines.py report.ncu-rep kernel_substring [top]"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "d2slam_b200", "libd2ba.so")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + ksub], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
si, ii, srci = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
sass = []
for r in rows[hi + 1:]:
    try:
        sass.append((int(r[0], 16), r[srci].strip(), int(r[si]), int(r[ii])))
    except Exception:
        pass
base = sass[0][0]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
    lines_of = {}
    for f in os.listdir(td):
        if not f.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, f)], capture_output=True, text=True).stdout
        cur_fn, cur_line, in_fn = None, None, False
        for ln in dis.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur_fn = m.group(1); in_fn = ksub in cur_fn; continue
            if not in_fn:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_line:
                lines_of[int(m.group(1), 16)] = cur_line
agg = collections.defaultdict(lambda: [0, 0])
tot = sum(s[2] for s in sass)
for addr, txt, smp, ins in sass:
    key = lines_of.get(addr - base, ("?", 0))
    agg[key][0] += smp; agg[key][1] += ins
src_cache = {}
def src(fn, ln):
    p = os.path.join(ROOT, "d2slam_b200", "csrc", fn)
    if p not in src_cache:
        try: src_cache[p] = open(p).read().splitlines()
        except Exception: src_cache[p] = []
    L = src_cache[p]
    return L[ln - 1].strip()[:100] if 0 < ln <= len(L) else ""
print(f"total samples {tot}, sass instrs {len(sass)}, mapped {len(lines_of)}")
for (fn, ln), (smp, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*smp/max(tot,1):5.1f}%  inst={ins:10d}  {fn}:{ln}  {src(fn, ln)}")
