"""Markdown summary of an ncu report (raw page) and of a launch list CSV -> stdout.

usage: python tools/ncu_summary.py full <report.ncu-rep | raw.csv>      |      python tools/ncu_summary.py list <launches.csv>
"""
import collections, csv, io, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "sm__inst_issued.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max",
]

if sys.argv[1] == "full":
    if sys.argv[2].endswith(".csv"):  # a raw page exported on the GPU box (ncu -i rep --page raw --csv)
        out = open(sys.argv[2]).read()
    else:
        out = subprocess.run(["ncu", "-i", sys.argv[2], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u, v = rows[0], rows[1], rows[2]
    kn = v[h.index("Kernel Name")]
    print(f"kernel: `{kn}`\n\n| metric | unit | value |\n|---|---|---|")
    for n in WANT:
        if n in h:
            i = h.index(n)
            print(f"| {n} | {u[i]} | {v[i]} |")
else:
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        d[r[ki].split("(")[0][:70]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    print("| kernel | launches | mean us | share of GPU time |\n|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot * 100:.2f} % |")
