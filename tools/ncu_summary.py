#!/usr/bin/env python
"""Condense an Nsight Compute report into the few counters DESIGN.md / bench.py quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name        # writes .csv and .json

The .json holds, per kernel launch in the report: duration, DRAM bytes read + written (the `traffic`
that bench.py's roofline object reports), tensor-pipe activity and L2 hit rate."""
import csv
import json
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "launch__cluster_dim_x", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(value.replace(",", "")) * scale


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    idx = [(w, head.index(w)) for w in WANT if w in head]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([n for n, _ in idx])
        w.writerow([units[i] for _, i in idx])
        for r in data:
            w.writerow([r[i] for _, i in idx])
    col = {n: i for n, i in idx}
    summary = []
    for r in data:
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        summary.append({"kernel": r[col["Kernel Name"]].split("(")[0], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
                        "duration_us": float(r[col["gpu__time_duration.sum"]].replace(",", "")),
                        "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
                        "tensor_pipe_active_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
                        "dram_throughput_pct_of_peak": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                        "achieved_dram_GBps": round((rd + wr) / float(r[col["gpu__time_duration.sum"]].replace(",", "")) / 1e3, 1),
                        "l2_hit_pct": float(r[col["lts__t_sector_hit_rate.pct"]])})
    json.dump({"report": rep, "command": "ncu --set full --clock-control none --import-source on", "launches": summary}, open(out + ".json", "w"), indent=1)
    for s in summary:
        print(s)


if __name__ == "__main__":
    main()
