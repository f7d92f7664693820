"""Aggregate an ncu `--page source --print-source cuda,sass --csv` export per CUDA source line."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < 10: continue
    if r[0] == "-" or not r[0].isdigit():
        continue
    # line summary row: has line number + source; samples column index 6, inst 7, thread inst 8
    try:
        key = (cur_file, int(r[0]), r[1].strip()[:110])
        samples = float(r[hdr["# Samples"]] or 0); inst = float(r[hdr["Instructions Executed"]] or 0); tinst = float(r[hdr["Thread Instructions Executed"]] or 0)
    except Exception:
        continue
    a = agg.setdefault(key, [0, 0, 0]); a[0] += samples; a[1] += inst; a[2] += tinst
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print(f"total samples {tot:.0f} total warp-inst {toti:.3e}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*v[0]/tot:5.1f}% smp {100*v[1]/toti:5.1f}% inst thr/inst {v[2]/max(v[1],1):5.1f} | {k[0]}:{k[1]} {k[2]}")
