"""Per-function, per-source-line totals from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: ncu_srclines.py file.csv <function substring> [topn]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 45
hdr = None; cur_file = None; func = None; line = None
agg = collections.OrderedDict()
def f(v):
    try: return float(v)
    except Exception: return 0.0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": func = r[1]; continue
    if len(r) > 10 and r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < 20 or func is None or want not in func: continue
    if r[0].isdigit(): line = (cur_file, int(r[0]), r[1].strip()[:90]); continue
    if r[2].startswith("0x"):
        a = agg.setdefault(line, [0.0] * 6)
        a[0] += f(r[hdr["# Samples"]]); a[1] += f(r[hdr["Instructions Executed"]]); a[2] += f(r[hdr["Thread Instructions Executed"]])
        a[3] += f(r[hdr["stall_no_inst"]]); a[4] += 1 if any(x in r[3] for x in (" DFMA", " DMUL", " DADD", " DSETP", " DMNMX")) else 0
        a[5] += f(r[hdr["Instructions Executed"]]) if any(x in r[3] for x in (" DFMA", " DMUL", " DADD", " DSETP", " DMNMX")) else 0
tot = [sum(v[i] for v in agg.values()) for i in range(6)]
print(f"{want}: samples {tot[0]:.0f} warp-inst {tot[1]:.3e} thr/inst {tot[2]/max(tot[1],1):.1f} no_inst samples {tot[3]:.0f} DP warp-inst {tot[5]:.3e}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*v[0]/tot[0]:5.1f}% smp {100*v[1]/tot[1]:5.1f}% inst {100*v[5]/max(tot[5],1):5.1f}% dp thr/inst {v[2]/max(v[1],1):5.1f} noinst {100*v[3]/max(v[0],1):4.0f}% | {k[0]}:{k[1]} {k[2]}")
