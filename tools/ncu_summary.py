"""Summarise an ncu report (raw page) into a small text file for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.txt [n_solves]"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    n_solves = int(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    lines = []
    summary = {}
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        lines.append(f"== kernel: {rec.get('Kernel Name')}  grid {rec.get('Grid Size')} block {rec.get('Block Size')}")
        for k in KEYS:
            if k in rec:
                lines.append(f"{k:75s} {rec[k]:>18s} {units[hdr.index(k)]}")
        stalls = [(float(rec[h].replace(',', '')), h[len(STALL):-len('_per_issue_active.ratio')]) for h in hdr
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h and rec[h]]
        lines.append("-- warp stall reasons (warps per issue-active cycle), top 8")
        for v, name in sorted(stalls, reverse=True)[:8]:
            lines.append(f"   {name:40s} {v:8.3f}")
        def num(k):
            u = units[hdr.index(k)]
            v = float(rec[k].replace(",", ""))
            return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
        dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        lines.append(f"dram bytes read+write per launch: {dram:.0f}")
        if n_solves:
            lines.append(f"dram bytes per solve: {dram / n_solves:.1f}  (n_solves = {n_solves})")
            summary = {"dram_bytes_per_launch": dram, "n_solves": n_solves, "dram_bytes_per_solve": dram / n_solves,
                       "source": rep}
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if summary:
        print(json.dumps(summary))


if __name__ == "__main__":
    main()
