"""Per-kernel DRAM traffic from an `ncu --set full` summary (tools/ncu_summary.py output): launches per step, mean
duration and DRAM bytes per launch. A kernel that is launched at several sizes inside one step (the radix pass:
eight passes over all keys, six over the low-support candidates) is reported for its LARGEST grid only, with the
other launches counted apart - `roofline.traffic` of the bench line is per full-size launch.

usage: traffic_json.py summary.csv n_keys out.json
"""
import csv
import json
import sys

src, n_keys, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(open(src)))
h = rows[0]
col = {name: h.index(name) for name in ("Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
                                        "dram__bytes_write.sum")}
units = rows[1]


def to_bytes(v, unit):
    f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * f[unit]


def to_ms(v, unit):
    f = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    return float(v) * f.get(unit, 1.0)


by = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").split("(")[0].split("<")[0].strip()
    grid = int(r[col["Grid Size"]].strip("()").split(",")[0])
    ms = to_ms(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
    b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
        to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    by.setdefault(name, []).append((grid, ms, b))
rec = {"_source": f"{src} (ncu --set full --clock-control none, cfg2 200 M reads, first step)", "_n_keys": n_keys}
n_steps = max(1, len(by.get("pass1_staged_kernel", [])))  # steps the capture covers
for name, ls in by.items():
    g = max(x[0] for x in ls)
    big = [x for x in ls if x[0] == g]
    rec[name] = {"launches_per_step": len(big) / n_steps, "ms_per_launch": sum(x[1] for x in big) / len(big),
                 "dram_bytes_per_launch": sum(x[2] for x in big) / len(big), "grid": g}
    if len(big) != len(ls):
        small = [x for x in ls if x[0] != g]
        rec[name]["smaller_launches"] = {"n": len(small), "ms_total": sum(x[1] for x in small),
                                         "dram_bytes_total": sum(x[2] for x in small)}
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in rec.items() if isinstance(v, dict) and v["ms_per_launch"] > 0.05}, indent=1))
