"""Turn the CSV of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum
--csv` pass into the per-launch table DESIGN.md cites (bench.py's `roofline.traffic` comes from the --set full capture,
profiles/ncu_summary.json).  usage: python profiles/launch_list.py launches.csv workload > list.txt"""
import csv, json, os, re, sys

path, workload = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
I = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    try:
        kid = int(r[I["ID"]])
    except ValueError:
        continue
    e = per.setdefault(kid, {"name": r[I["Kernel Name"]]})
    v = float(r[I["Metric Value"]].replace(",", ""))
    unit = r[I["Metric Unit"]]
    m = r[I["Metric Name"]]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit) if m.startswith("gpu__time") else None
    if m.startswith("gpu__time"):
        e["ms"] = v * scale
    elif m.startswith("dram__bytes"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
        e["rd" if "read" in m else "wr"] = v * mult
    elif m.startswith("smsp__inst_executed"):
        e["inst"] = v
ids = sorted(per)
# one render = from one k_prepare_scene to the next
starts = [i for i in ids if per[i]["name"].startswith("k_prepare_scene")]
first = starts[0] if starts else ids[0]
last = starts[1] if len(starts) > 1 else ids[-1] + 1
frame = [i for i in ids if first <= i < last]
total = sum(per[i].get("ms", 0.0) for i in frame)
print(f"# {workload} on one B200: every kernel of one render, ncu --metrics gpu__time_duration.sum,dram__bytes_*,smsp__inst_executed.sum")
print("# --clock-control none (kernels serialised by ncu: compare SHARES of the frame, not absolutes; the live frame overlaps stage A and stage B)")
print(f"# raw: {path}")
print(f"{'id':>3} {'kernel':60s} {'ms':>8} {'share':>6} {'Ginst':>8} {'GB read':>8} {'GB written':>10}")
for i in frame:
    e = per[i]
    name = re.sub(r"atmrt::", "", e["name"])
    print(f"{i:3d} {name[:60]:60s} {e.get('ms', 0):8.3f} {100 * e.get('ms', 0) / total:5.1f}% {e.get('inst', 0) / 1e9:8.3f} {e.get('rd', 0) / 1e9:8.2f} {e.get('wr', 0) / 1e9:10.2f}")
print(f"sum of one render: {total:.3f} ms (serialised)")
