"""Per-kernel shares from an ncu launch list (`--metrics gpu__time_duration.sum --clock-control none --csv`).
usage: ncu_launch_list.py list.csv out.md "<title line>" """
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", "")[:90]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[14]) / 1e6
tot = sum(v[1] for v in agg.values())
with open(sys.argv[2], "w") as f:
    f.write(f"# {sys.argv[3]}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES). Raw CSV: {sys.argv[1].split('/')[-1]}\n\n")
    f.write(f"total profiled device time {tot:.1f} ms over {len(rows)} launches\n\n| kernel | launches | total ms | share | ms per launch |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.3f} |\n")
print("wrote", sys.argv[2], f"{tot:.1f} ms", len(rows))
