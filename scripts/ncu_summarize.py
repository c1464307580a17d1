"""Turn an `ncu --set full` report into the two JSON files kept under profiles/:

    python scripts/ncu_summarize.py gpurun_out/prof_r1.ncu-rep profiles/r1_ncu_full_summary_4096chains.json profiles/r1_traffic.json

The summary lists every captured launch (duration, DRAM bytes, DRAM / tensor / L2 / SM utilisation, registers);
the traffic file averages `dram__bytes_read.sum + dram__bytes_write.sum` over the whole-MLP launches of one round,
which is what bench.py reports as `roofline.traffic`.  Runs where the report was read (ncu -i), not on the GPU box.
"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main(report, summary_path, traffic_path=None):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    header, units = rows[0], rows[1]

    def value(row, key):
        if key not in header:
            return None
        i = header.index(key)
        try:
            return float(row[i]) * UNIT_SCALE.get(units[i], 1.0)
        except ValueError:
            return None

    launches = []
    for row in rows[2:]:
        entry = {"kernel": row[header.index("Kernel Name")][:80], "grid": row[header.index("Grid Size")],
                 "block": row[header.index("Block Size")]}
        for key in KEYS:
            entry[key] = value(row, key)
        entry["duration_us"] = entry["gpu__time_duration.sum"]
        entry["dram_bytes"] = (entry["dram__bytes_read.sum"] or 0.0) + (entry["dram__bytes_write.sum"] or 0.0)
        launches.append(entry)
    with open(summary_path, "w") as f:
        json.dump(launches, f, indent=1)
    if traffic_path:
        out = {}
        for key, pattern in (("fused_mlp_kernel", "fused_mlp"), ("x3_mlp_kernel", "x3_mlp")):   # bf16 / fp32-accurate whole-MLP kernels
            mlp = [e for e in launches if pattern in e["kernel"]]
            if mlp:
                out[key] = {
                    "dram_bytes_per_launch": sum(e["dram_bytes"] for e in mlp) / len(mlp), "launches_captured": len(mlp),
                    "per_launch": [{"duration_us": e["duration_us"], "dram_bytes": e["dram_bytes"],
                                    "tensor_active_pct": e["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]}
                                   for e in mlp]}
        for name in ("clause_gather", "literal_gather", "pairnorm", "spmm_rows"):
            sel = [e for e in launches if name in e["kernel"]]
            if sel:
                out[name] = {"dram_bytes_per_launch": sum(e["dram_bytes"] for e in sel) / len(sel),
                             "duration_us": sum(e["duration_us"] for e in sel) / len(sel), "launches_captured": len(sel)}
        with open(traffic_path, "w") as f:
            json.dump(out, f, indent=1)
    print("%d launches summarised" % len(launches))


if __name__ == "__main__":
    main(*sys.argv[1:4])
