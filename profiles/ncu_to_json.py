"""Merge an `ncu --set full --import-source on` capture into profiles/ncu_summary.json: per kernel the counters DESIGN.md
and bench.py's `roofline` cite -- duration, issue-active %, FP64-pipe %, warp instructions (all, and those of the FP64
pipe counted on the SASS page), DRAM bytes, occupancy, cache hit rates, top stall reasons -- keyed "<workload>:<kernel>",
with the sha of the sources the capture was taken from (profiles/source_sha.py, written on the GPU box next to the report).
usage: python profiles/ncu_to_json.py REPORT.ncu-rep WORKLOAD SHA [NOTE]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, workload, sha = sys.argv[1], sys.argv[2], sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ""
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ncu_summary.json")


def page(*args):
    return subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(page("--page", "raw"))))
hdr, units = rows[0], rows[1]
I = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = r[I[name]].replace(",", "")
    u = units[I[name]]
    x = float(v) if v not in ("", "n/a") else float("nan")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
    return x * scale.get(u, 1.0)


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)", name)
    return m.group(1) if m else name


summary = json.load(open(OUT)) if os.path.exists(OUT) else {}
stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
for r in rows[2:]:
    kid = r[I["ID"]]
    name = short(r[I["Kernel Name"]])
    # FP64-pipe warp instructions: the D* opcodes of the SASS page of this launch
    fp64 = total = 0
    seen = set()
    src = list(csv.reader(io.StringIO(page("--page", "source", "--print-source", "sass", "--kernel-name", "regex:" + name + r"\b"))))
    sh = next((x for x in src if x and x[0] == "Address"), None)
    if sh:
        iS, iN = sh.index("Source"), sh.index("Instructions Executed")
        for x in src:
            if len(x) <= iN or not x[iN].isdigit() or x[0] in seen:
                continue
            seen.add(x[0])
            ops = x[iS].strip().split()
            op = ops[1] if ops and ops[0].startswith("@") and len(ops) > 1 else (ops[0] if ops else "")
            total += int(x[iN])
            if re.match(r"D(ADD|MUL|FMA|SETP|MNMX)\b", op.split(".")[0]):
                fp64 += int(x[iN])
    stalls = sorted(((float(r[I[h]] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall_cols), reverse=True)[:4]
    summary[f"{workload}:{name}"] = {
        "gpu_time_ms": val(r, "gpu__time_duration.sum"),
        "grid": r[I["launch__grid_size"]], "block": r[I["launch__block_size"]], "registers": int(float(r[I["launch__registers_per_thread"]])),
        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_pct": val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "warp_instructions": val(r, "smsp__inst_executed.sum"),
        "fp64_warp_instructions": fp64, "sass_warp_instructions": total,
        "threads_per_instruction": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
        "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
        "top_stalls_warps_per_issue_cycle": {k: round(v, 2) for v, k in stalls},
        "source": f"{os.path.basename(rep)} (ncu --set full --clock-control none, kernel alone){' -- ' + note if note else ''}",
        "source_sha": sha,
    }
    print(name, summary[f"{workload}:{name}"]["gpu_time_ms"], "ms, fp64 warp instr", fp64, "of", total)
json.dump(summary, open(OUT, "w"), indent=1)
