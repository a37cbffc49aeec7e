"""Summarises `ncu --set full` reports: the metrics the roofline discussion uses -> profiles/<name>.txt, and the DRAM
bytes per launch -> profiles/ncu_traffic.json (read by bench.py for roofline.traffic).
usage: python tools/ncu_summary.py <report.ncu-rep> <profiles/out.txt> [<op>:<shape key for ncu_traffic.json>]"""
import csv
import json
import subprocess
import sys
from pathlib import Path

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], Path(sys.argv[2])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = [f"# ncu --set full --clock-control none: {d['Kernel Name'][0]}", f"# report: {Path(rep).name}"]
    for k in KEYS:
        if k in d:
            lines.append(f"{k} = {d[k][0]} {d[k][1]}")
    def num(k):
        v, u = d[k]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    dur_us = float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(d["gpu__time_duration.sum"][1], 1)
    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    lines.append(f"derived: dram read+write = {dram / 1e6:.2f} MB per launch, {dram / dur_us / 1e3:.1f} GB/s achieved under ncu "
                 f"(of the measured 6545 GB/s copy peak: {dram / dur_us / 1e3 / 6545:.3f})")
    out.write_text("\n".join(lines) + "\n")
    print("\n".join(lines))
    if len(sys.argv) > 3:
        tp = out.parent / "ncu_traffic.json"
        db = json.loads(tp.read_text()) if tp.exists() else {}

        db[sys.argv[3]] = {
            "dram_read_bytes": num("dram__bytes_read.sum"), "dram_write_bytes": num("dram__bytes_write.sum"),
            "duration_us_under_ncu": dur_us, "dram_gbs_under_ncu": dram / dur_us / 1e3,
            "tensor_pipe_active_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]),
            "l2_to_sm_bytes": num("l1tex__m_xbar2l1tex_read_bytes.sum") if "l1tex__m_xbar2l1tex_read_bytes.sum" in d else None,
            "report": str(out),
        }
        tp.write_text(json.dumps(db, indent=1))


if __name__ == "__main__":
    main()
