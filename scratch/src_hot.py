"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu -i rep --page source --print-source cuda,sass --csv --kernel-name regex:K | python src_hot.py [top]"""
import csv, sys, collections
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = csv.reader(sys.stdin)
cur_file = None; hdr = None
agg = collections.OrderedDict()
tot_s = tot_i = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed"); continue
    if hdr is None: continue
    if r[0] != "":  # a source line header row (aggregated over its SASS)
        key = (cur_file, int(r[0]))
        f = lambda v: int(v) if v.isdigit() else 0
        s, i, t = f(r[iS]), f(r[iI]), f(r[iT])
        if key in agg:
            a = agg[key]; a[0] += s; a[1] += i; a[2] += t
        else:
            agg[key] = [s, i, t, r[1].strip()[:110]]
        tot_s += s; tot_i += i
print(f"total samples {tot_s}, warp instructions {tot_i}")
for (f, ln), (s, i, t, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*i/max(tot_i,1):5.1f}% ins thr/ins {t/max(i,1):4.1f}  {f}:{ln}  {src}")
