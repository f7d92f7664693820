#!/bin/bash
# build_ref_oracle.sh - build the REAL reference as a parity oracle when a Fortran compiler exists.
#
# Nothing of /root/reference is copied into the repository: the module text (everything before `program run_sph`,
# SUMMER_SPH.f90:934 | "SUMMER_SPH - Variable.f90":1168) is extracted with sed into oracle/_ref/build/ (git-ignored) at
# build time and linked with this repository's own dump drivers (oracle/ref_driver/ref_dump_{f,v}.f90), at -O0 (the
# README's build line has no -O) and -O3, without -fopenmp (the OpenMP pair loop of the reference is racy, F:302-313)
# and without -march=native (no FMA contraction).  Outputs: oracle/_ref/ref_dump_{f,v}_{O0,O3}.
# Without a Fortran compiler (this image and the GPU boxes have none: profiles/r2_fortran_probe_gpubox.log) it says so
# and exits 0; tests/test_oracle_vs_ref.py then skips and the oracle stays "parity unpinned".
set -u
cd "$(dirname "$0")/.."
REF=${SPH_REFERENCE_DIR:-/root/reference}
FC=""
for c in gfortran flang-new flang nvfortran ifx ifort lfortran; do
  if command -v "$c" >/dev/null 2>&1; then FC=$c; break; fi
done
if [ -z "$FC" ]; then echo "build_ref_oracle: no Fortran compiler on this machine (looked for gfortran flang nvfortran ifx ifort lfortran) - nothing built"; exit 0; fi
if [ ! -f "$REF/SUMMER_SPH.f90" ]; then echo "build_ref_oracle: $REF not present - nothing built"; exit 0; fi
mkdir -p oracle/_ref/build
sed '/^program run_sph/,$d' "$REF/SUMMER_SPH.f90" > oracle/_ref/build/module_f.f90
sed '/^program run_sph/,$d' "$REF/SUMMER_SPH - Variable.f90" > oracle/_ref/build/module_v.f90
rc=0
for opt in O0 O3; do
  for v in f v; do
    d=oracle/_ref/build/${v}_$opt; mkdir -p "$d"
    ( cd "$d" && $FC -$opt -o ../../ref_dump_${v}_$opt ../module_$v.f90 ../../../ref_driver/ref_dump_$v.f90 ) || { echo "build_ref_oracle: $FC failed for $v -$opt"; rc=1; }
  done
done
# the Fortran host of this repository (host/run_sph_b200.f90) needs the same compiler
if [ -f summersph_b200/libsph_b200.so ]; then make -C host run_sph_b200 FC=$FC || rc=1; fi
ls -la oracle/_ref/ 2>/dev/null
exit $rc
