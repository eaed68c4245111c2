#!/usr/bin/env python3
"""Summarises an ncu capture into a short text file for profiles/.  Input: an .ncu-rep, or the CSV that
`ncu -i x.ncu-rep --page raw --csv` printed on the GPU box (tools/capture_profiles.sh exports that, the reports themselves are too big).
    python tools/ncu_summary.py gpurun_out/r01d_ntt_raw.csv profiles/r01_ncu_ntt_pass_kernel.txt [traffic.json]
With a third argument also writes the per-launch DRAM traffic (bench.py's roofline.traffic) as JSON."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_issue_stalled",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum", "derived__smsp__inst_executed_op_local_ld.sum"]


def to_bytes(value, unit):
    v = float(value)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    lines = [f"# summary of {rep} (ncu --set full --clock-control none --import-source on; python tools/prove_once.py 20 2)"]
    traffic = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append(f"\n== launch id {r[0]}: {name}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k:
                    lines.append(f"{k:90s} {r[i]:>18s} {units[i]}")
        tot, d = 0.0, {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued"):
                try:
                    d[h[33:]] = float(r[i]); tot += float(r[i])
                except ValueError:
                    pass
        if tot:
            lines.append("stall samples: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(d.items(), key=lambda x: -x[1])[:10]))
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        traffic.append({"kernel": name, "grid": r[hdr.index("Grid Size")], "dram_read_bytes": to_bytes(r[rd], units[rd]),
                        "dram_write_bytes": to_bytes(r[wr], units[wr]), "duration_ms": float(r[hdr.index("gpu__time_duration.sum")])})
    if len(sys.argv) > 3:
        import json
        per_launch = sum(t["dram_read_bytes"] + t["dram_write_bytes"] for t in traffic) / len(traffic)
        json.dump({"source": rep, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full", "launches": traffic,
                   "per_launch_bytes": per_launch}, open(sys.argv[3], "w"), indent=1)
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
