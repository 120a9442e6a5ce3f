"""Turn the raw page of an `ncu --set full` capture of the block kernels into an entry of
profiles/ncu_traffic.json, which bench.py reads for `roofline.traffic` (DRAM bytes per launch of
the dominant kernel, for the same window and block size).

    ncu -i gpurun_out/<capture>.ncu-rep --page raw --csv > raw.csv
    python scripts/ncu_traffic.py raw.csv <window_bytes> <block_bytes> profiles/<copy of raw.csv>
"""
import csv
import json
import os
import re
import sys

UNIT_FACTORS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
                "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
MODES = {0: "K1_fill_pattern", 1: "K2_verify_pattern", 2: "K3_fill_random_pct100"}


def main():
    raw_path, window_bytes, block_bytes, source = sys.argv[1], int(sys.argv[2]), \
        int(sys.argv[3]), sys.argv[4]
    rows = list(csv.reader(open(raw_path)))
    header, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(header)}

    def value(row, name):
        return float(row[col[name]].replace(",", "")) * UNIT_FACTORS.get(units[col[name]], 1)

    kernels = {}
    for row in rows[2:]:
        name = row[col["Kernel Name"]]
        match = re.search(r"elb_blocks_\w*kernel<(?:\(int\))?\s*(\d+)(?:,\s*(?:\(int\))?\s*(\d+))?>", name)
        if not match or (match.group(2) not in (None, "0")):
            continue  # (only the resident forms: STAGE_NONE)
        key = MODES.get(int(match.group(1)))
        if key is None or key in kernels:
            continue
        kernels[key] = {
            "kernel": name.split("(")[0].replace("void ", ""),
            "dram_bytes_read": int(value(row, "dram__bytes_read.sum")),
            "dram_bytes_write": int(value(row, "dram__bytes_write.sum")),
            "duration_us": round(value(row, "gpu__time_duration.sum"), 2),
            "registers_per_thread": int(value(row, "launch__registers_per_thread")),
            "dram_throughput_pct": round(value(row, "dram__throughput.avg.pct_of_peak_sustained_elapsed"), 1)
            if "dram__throughput.avg.pct_of_peak_sustained_elapsed" in col else None,
            "sm_throughput_pct": round(value(row, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), 1)
            if "sm__throughput.avg.pct_of_peak_sustained_elapsed" in col else None,
        }
    entry = {"window_bytes": window_bytes, "block_bytes": block_bytes, "source": source,
             "tool": "ncu --set full --clock-control none (one launch per kernel)",
             "kernels": kernels}
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                            "profiles", "ncu_traffic.json")
    try:
        data = json.load(open(out_path))
    except Exception:
        data = {"captures": []}
    data["captures"] = [c for c in data.get("captures", [])
                        if (c["window_bytes"], c.get("block_bytes")) != (window_bytes, block_bytes)]
    data["captures"].append(entry)
    with open(out_path, "w") as f:
        json.dump(data, f, indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
