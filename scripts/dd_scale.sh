#!/bin/bash
# Developer check of the Morton-domain form at benchmark-like sizes with VIRTUAL ranks (host-segment collectives, all
# ranks on cuda:0): domain statistics + the global fingerprint sums against the 1-rank run of the same disc.
#   scripts/dd_scale.sh TAG N STEPS WORLD [WORLD ...]
TAG=$1; N=$2; STEPS=$3; shift 3
mkdir -p gpurun_out
for W in "$@"; do
  TOKEN="/sphb200_dd_${TAG}_${W}_$$"
  pids=()
  for ((r = 0; r < W; r++)); do
    timeout 600 python scripts/dd_scale_rank.py $r $W host 0 $N $STEPS $TOKEN gpurun_out/${TAG}_n${N}_w${W}_r${r}.json > gpurun_out/${TAG}_n${N}_w${W}_r${r}.log 2>&1 &
    pids+=($!)
  done
  rc=0
  for p in "${pids[@]}"; do wait $p || rc=$?; done
  echo "== $TAG N=$N world=$W rc=$rc"
  python - <<EOF
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_n${N}_w${W}_r*.json")):
    d = json.load(open(f))
    print(d["rank"], "sizes", d["sizes"], "dt", d["dt"], "domain", d["domain"], "sums", d["sums"])
    print("   last step", d["stages"][-1])
    print("   counters", d["counters"])
EOF
  tail -3 gpurun_out/${TAG}_n${N}_w${W}_r0.log | cut -c1-400
done
