"""profiles/r2_traffic.json from an `ncu --set full` capture of tools/prof_targets.py: per key, the DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum) and duration of its kernels, per launch.

    python tools/ncu_traffic.py gpurun_out/r2_prof.ncu-rep gpurun_out/r2_prof_order.json > profiles/r2_traffic.json"""
import csv
import json
import subprocess
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}


def main(rep, order_path):
    order = json.load(open(order_path))
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                      "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
                                      "sm__warps_active.avg.pct_of_peak_sustained_active")}
    launches = rows[2:]
    if len(launches) != len(order):
        print("warning: %d profiled launches vs %d expected" % (len(launches), len(order)), file=sys.stderr)
    res = {}
    for (key, sub), row in zip(order, launches):
        name = row[col["Kernel Name"]]
        if sub not in name:
            print("warning: launch order mismatch: expected *%s*, got %s" % (sub, name[:60]), file=sys.stderr)
        if key == "warmup":
            continue
        rd = float(row[col["dram__bytes_read.sum"]].replace(",", "")) * UNIT.get(units[col["dram__bytes_read.sum"]], 1)
        wr = float(row[col["dram__bytes_write.sum"]].replace(",", "")) * UNIT.get(units[col["dram__bytes_write.sum"]], 1)
        us = float(row[col["gpu__time_duration.sum"]].replace(",", "")) * TIME.get(units[col["gpu__time_duration.sum"]], 1)
        e = res.setdefault(key, {"kernels": [], "traffic_bytes_per_launch": 0, "duration_us_under_ncu": 0.0})
        e["kernels"].append({"name": name.split("(")[0][:100], "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                             "duration_us": round(us, 2),
                             "dram_throughput_pct": float(row[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                             "registers": int(float(row[col["launch__registers_per_thread"]])),
                             "warps_active_pct": float(row[col["sm__warps_active.avg.pct_of_peak_sustained_active"]])})
        e["traffic_bytes_per_launch"] += int(rd + wr)
        e["duration_us_under_ncu"] = round(e["duration_us_under_ncu"] + us, 2)
    res["_source"] = "ncu --set full --clock-control none of tools/prof_targets.py (round 2); per-kernel cache flush: cold"
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
