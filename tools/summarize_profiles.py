"""Turn the ncu captures under gpurun_out/ into committed text summaries under profiles/.
usage: summarize_profiles.py <tag> [B]"""
import csv, os, re, subprocess, sys, json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
B = sys.argv[2] if len(sys.argv) > 2 else "256"
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_dmma.sum.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks", "sm__maximum_warps_per_active_cycle_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
launch = os.path.join(G, f"launches_B{B}.csv")
out = []
if os.path.exists(launch):
    rows = [l for l in open(launch) if l.startswith('"')]
    r = list(csv.reader(rows)); h = r[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
    agg = {}
    for row in r[1:]:
        n = re.sub(r"\(.*", "", row[ki]); v = float(row[vi].replace(",", ""))
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v for _, v in agg.values())
    out.append(f"# launch list (ncu --metrics gpu__time_duration.sum --clock-control none), batch {B} W1 windows, non-graph solve\n")
    out.append("# cold-cache / serialised per-launch times: compare SHARES\n")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k:40s} launches={n:3d} avg={v/n/1e3:9.2f} us share={100*v/tot:5.1f}%\n")
    open(os.path.join(P, f"{tag}_launches_B{B}.txt"), "w").writelines(out)
    import shutil; shutil.copy(launch, os.path.join(P, f"{tag}_launches_B{B}.csv"))
traffic = {}
for f in sorted(os.listdir(G)):
    m = re.match(rf"prof_(\w+)_(B{B}|A8)\.ncu-rep", f)
    if not m or m.group(1) in ("multi", "chol2"):
        continue
    kn = m.group(1); sfx = m.group(2)
    raw = subprocess.run(["ncu", "-i", os.path.join(G, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    if len(r) < 3:
        continue
    h, u, v = r[0], r[1], r[-1]
    what = f"batch {B} W1 windows" if sfx.startswith("B") else "74 eight-agent swarms (592 windows) on one GPU"
    lines = [f"# ncu --set full --clock-control none --import-source on, kernel k_{kn}, {what} ({f})\n"]
    vals = {}
    for i, n in enumerate(h):
        if n in KEYS:
            lines.append(f"{n:90s} {v[i]:>16s} {u[i]}\n"); vals[n] = (v[i], u[i])
    def tobytes(key):
        if key not in vals: return None
        x, unit = vals[key]; x = float(x.replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    rd, wr = tobytes("dram__bytes_read.sum"), tobytes("dram__bytes_write.sum")
    if rd is not None and wr is not None:
        if sfx.startswith("B"):
            traffic[kn] = rd + wr
        lines.append(f"dram traffic per launch (read+write): {(rd+wr)/1e6:.1f} MB\n")
    src = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), os.path.join(G, f), f"k_{kn}", "16"], capture_output=True, text=True).stdout
    lines.append("\n# warp-stall samples by CUDA source line (top 16)\n" + src)
    open(os.path.join(P, f"{tag}_ncu_{kn}_{sfx}.txt"), "w").writelines(lines)
if traffic:
    json.dump({f"{k}_bytes_per_launch": v for k, v in traffic.items()} | {"batch": int(B), "tag": tag}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print("wrote", sorted(os.listdir(P)))
