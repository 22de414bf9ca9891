#!/bin/bash
# Dev tool, run on the GPU box: time tools/pair_ab.py for library variants (tools/build_variants.sh) and / or
# environment settings.  usage: tools/ab_run.sh [lib:<variant> | env:<label>:<VAR=VALUE,...>] ...
mkdir -p gpurun_out
cp mnemophonix_b200/libmnemo_cuda.so /tmp/lib_default.so
for spec in "$@"; do
  kind=${spec%%:*}; rest=${spec#*:}
  if [ "$kind" = lib ]; then
    cp mnemophonix_b200/_variants/libmnemo_cuda.so.$rest mnemophonix_b200/libmnemo_cuda.so
    timeout 200 python tools/pair_ab.py $rest >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err
    cp /tmp/lib_default.so mnemophonix_b200/libmnemo_cuda.so
  else
    label=${rest%%:*}; vars=${rest#*:}
    env $(echo $vars | tr ',' ' ') timeout 200 python tools/pair_ab.py $label >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err
  fi
done
cat gpurun_out/ab.jsonl
