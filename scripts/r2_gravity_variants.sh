#!/bin/bash
# Round-2 helper for the unmeasured experiments of DESIGN.md §9: GW_SUBLISTS (k_gravity's interaction list read through
# per-lane-group index lists), GROUP_SPLIT_BUCKETS (walk groups cut into full 32-particle runs inside a run of sibling buckets)
# and GW_FAR_REUSE (evaluation A keeps evaluation B's far-field gravity and re-walks only what can hold a term with dist < 2h).
#   here (no GPU):   scripts/r2_gravity_variants.sh build              # nvcc -> summersph_b200/variants/libsph_{sub2,sub4,sub8,split,split_sub4,far}.so
#   on the GPU box:  gpurun --timeout 900 -- 'scripts/r2_gravity_variants.sh run 16e6 > gpurun_out/r2_variants.log 2>&1'
# `run` first holds every variant to the oracle (the parity tests through the C-ABI, SPH_B200_LIB selects the
# library), then times the default library and the variants at N particles (per-stage device times of 3 steps).
set -u
cd "$(dirname "$0")/.."
case "${1:-}" in
  build)
    for q in 2 4 8; do scripts/build_variant.sh sub$q "-DGW_SUBLISTS=$q" || exit 1; done
    scripts/build_variant.sh split "-DGROUP_SPLIT_BUCKETS" || exit 1
    scripts/build_variant.sh split_sub4 "-DGROUP_SPLIT_BUCKETS -DGW_SUBLISTS=4" || exit 1
    scripts/build_variant.sh far "-DGW_FAR_REUSE" || exit 1 ;;
  run)
    N=${2:-16e6}
    for v in sub2 sub4 sub8 split split_sub4 far; do
      lib=summersph_b200/variants/libsph_$v.so
      [ -f "$lib" ] || { echo "missing $lib (run 'build' before gpurun)"; continue; }
      echo "== parity $v"
      # far: evaluation A adds the same gravity terms in another order, so the "bit-identical to a rebuild" test does not apply
      SPH_B200_LIB=$lib SPH_B200_FAR_STATS=1 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not tree_reuse_is_bit_identical" 2>&1 | tail -5
    done
    scripts/gpu_variants.sh "$N" 3 sub2 sub4 sub8 split split_sub4 far ;;
  *) echo "usage: $0 build | run [N]"; exit 2 ;;
esac
