mkdir -p gpurun_out; rm -f gpurun_out/cross.jsonl
for k in 1250 2500 5000; do
  for m in 0 1; do
    MNEMO_PAIR_INDEX=$m timeout 200 python tools/pair_ab.py "shard$k-pairs$m" $k >> gpurun_out/cross.jsonl 2>> gpurun_out/cross.err
  done
done
cat gpurun_out/cross.jsonl
