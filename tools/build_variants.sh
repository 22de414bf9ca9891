#!/bin/bash
# Dev tool: build variants of libmnemo_cuda.so that differ in the -D flags of ONE source file, for A/B timing in one
# gpurun call.  usage: tools/build_variants.sh k_pairprobe.cu "name1:-DX=1 -DY=0" "name2:-DX=0" ...
# Each variant lands in mnemophonix_b200/_variants/libmnemo_cuda.so.<name>; on the box: cp it over mnemophonix_b200/libmnemo_cuda.so.
set -e
cd "$(dirname "$0")/.."
SRC=$1; shift
python -m mnemophonix_b200.build > /dev/null
mkdir -p mnemophonix_b200/_variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ftz=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC,-O2,-fno-fast-math,-ffp-contract=off -Xptxas -v"
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  obj=/tmp/variant_${name}.o
  nvcc $FLAGS $defs -c mnemophonix_b200/csrc/$SRC -o $obj 2> mnemophonix_b200/_variants/${name}.ptxas.txt
  objs=$(ls mnemophonix_b200/_obj/*.o | grep -v "/${SRC%.cu}.o")
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o mnemophonix_b200/_variants/libmnemo_cuda.so.${name} $objs $obj -cudart static
  echo "$name: $(grep -A2 'search_pair_kernel' mnemophonix_b200/_variants/${name}.ptxas.txt | grep Used | head -1)"
done
