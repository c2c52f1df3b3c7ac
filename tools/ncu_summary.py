"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed) into profiles/<name>.txt and update
profiles/r2_traffic.json (per-launch DRAM bytes of the bench's roofline kernel, read by bench.py).

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r1_x_ncu.txt "header text" [--traffic-key KEY]
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__inst_executed.avg.per_cycle_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")


def to_bytes(v, unit):
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v.replace(",", "")) * mul


def main():
    rep, out, header = sys.argv[1], sys.argv[2], sys.argv[3]
    key = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# {header}", f"# source: {os.path.basename(rep)} -- ncu --set full --clock-control none --import-source on (gpurun, one B200, cold cache, serialised); per launch", ""]
    traffic = None
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        lines.append(f"{'Kernel Name':85s} {d['Kernel Name'][0]}")
        for h in sorted(d):
            if h in KEEP or ("issue_stalled" in h and "per_issue_active" in h):
                lines.append(f"{h:85s} {d[h][0]:>20s} {d[h][1]}")
        lines.append("")
        traffic = {"kernel": d["Kernel Name"][0], "dram_bytes_read": to_bytes(*d["dram__bytes_read.sum"]),
                   "dram_bytes_write": to_bytes(*d["dram__bytes_write.sum"]), "ncu_report": os.path.relpath(out, os.path.dirname(os.path.abspath(__file__)) + "/..")}
    open(out, "w").write("\n".join(lines))
    if key and traffic:
        path = os.path.join(os.path.dirname(out), os.environ.get("B200Q_TRAFFIC_JSON", "r2_traffic.json"))
        tj = json.load(open(path)) if os.path.exists(path) else {}
        tj[key] = traffic
        json.dump(tj, open(path, "w"), indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
