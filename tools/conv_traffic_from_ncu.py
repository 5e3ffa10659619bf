"""profiles/conv_traffic.json from `ncu --set full` captures of the dominant convolution (one per precision mode):
    python tools/conv_traffic_from_ncu.py <precision>=<raw.csv> ...     (raw.csv = `ncu -i X.ncu-rep --page raw --csv`)
bench.py reads roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) from that file."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out_path = os.path.join(ROOT, "profiles", "conv_traffic.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
key = "conv3d_64x64_64x64x64"
for arg in sys.argv[1:]:
    prec, path = arg.split("=")
    rows = [r for r in csv.reader(open(path)) if r]
    hdr, units, row = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(name):
        return float(row[idx[name]].replace(",", "")) * UNIT[units[idx[name]]]
    data.setdefault(key, {})[prec] = {
        "batch": 8, "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
        "gpu_time_us": float(row[idx["gpu__time_duration.sum"]].replace(",", "")) * {"usecond": 1, "us": 1, "msecond": 1e3, "ms": 1e3, "nsecond": 1e-3, "ns": 1e-3}[units[idx["gpu__time_duration.sum"]]],
        "tensor_active_pct": float(row[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
        "kernel": row[idx["Kernel Name"]],
        "source": f"ncu --set full --clock-control none, profiles/{os.path.basename(path).replace('_raw.csv', '')}_ncu_full.txt (one launch, B = 8, round 2 HEAD)"}
json.dump(data, open(out_path, "w"), indent=1)
print(json.dumps(data, indent=1))
