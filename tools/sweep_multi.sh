#!/bin/bash
# BASELINE.json configs[4] on one box: bench.py at N in {1,2,4,8} x batch 2^14..2^22 (weak: per GPU; strong: in total), D1 (+ D2 at
# three sizes), resident inputs only (--no-e2e --no-cpu-baseline).  One JSON line per run -> $OUT.
OUT=${1:-gpurun_out/r02_sweep.jsonl}
NS=${2:-"1 2 4 8"}
mkdir -p $(dirname $OUT); : > $OUT
run() {  # N, extra args
  local n=$1; shift
  if [ "$n" = 1 ]; then timeout 200 python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e --no-cpu-baseline "$@" 2>/dev/null >> $OUT
  else timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 100 --warmup 5 --no-e2e --no-cpu-baseline "$@" 2>/dev/null >> $OUT
  fi
}
for n in $NS; do
  for lb in 14 16 18 20 22; do run $n --log2-batch $lb --no-other-configs; done
  for lb in 20 22; do run $n --log2-batch $lb --strong --no-other-configs; done
  for lb in 16 20; do run $n --log2-batch $lb --dist D2 --no-other-configs; done
  run $n --log2-batch 20                      # with the cart / pendulum shards (other_configs)
done
wc -l $OUT
