"""Summarise an .ncu-rep (first profiled kernel) into the handful of numbers DESIGN.md quotes.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "sass__inst_executed_global_loads",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out = [f"# ncu summary of {rep}", f"kernel: {d.get('Kernel Name', ('?', ''))[0]}"]
    for k in KEYS:
        if k in d:
            out.append(f"{k}: {d[k][0]} {d[k][1]}")
    stalls = [(h, float(v)) for h, (v, u) in d.items()
              if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and v]
    out.append("stall reasons (warps per issue):")
    for h, v in sorted(stalls, key=lambda x: -x[1])[:10]:
        out.append(f"  {v:6.2f} " + h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
