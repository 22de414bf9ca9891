#!/bin/bash
# Checked build of libmnemo_cuda.so: every .cu with -DMNEMO_CHECKED (csrc/common.cuh: MN_CHECK bounds assertions on the
# data-dependent indices of the search kernels trap on violation).  Output: mnemophonix_b200/_checked/libmnemo_cuda.so.
#   tools/checked_build.sh                 (here: nvcc cross-compiles)
#   tools/checked_build.sh run [pytest args]   (on the GPU box: swap the library in, run the GPU tests, swap it back)
set -e
cd "$(dirname "$0")/.."
OUT=mnemophonix_b200/_checked
if [ "$1" = run ]; then
  shift
  cp mnemophonix_b200/libmnemo_cuda.so /tmp/libmnemo_cuda.so.plain
  cp $OUT/libmnemo_cuda.so mnemophonix_b200/libmnemo_cuda.so
  set +e
  python -m pytest tests -m gpu -q "$@"
  rc=$?
  cp /tmp/libmnemo_cuda.so.plain mnemophonix_b200/libmnemo_cuda.so
  exit $rc
fi
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ftz=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC,-O2,-fno-fast-math,-ffp-contract=off -DMNEMO_CHECKED"
objs=""
for f in mnemophonix_b200/csrc/*.cu; do
  o=$OUT/$(basename ${f%.cu}).o
  nvcc $FLAGS -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libmnemo_cuda.so $objs -cudart static
rm -f $OUT/*.o
echo "built $OUT/libmnemo_cuda.so ($(grep -c MN_CHECK mnemophonix_b200/csrc/k_pairprobe.cu mnemophonix_b200/csrc/k_search.cu | tr '\n' ' ')assertions)"
