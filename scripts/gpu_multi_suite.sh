#!/bin/bash
# multi-GPU evidence: the contract line + BASELINE configs 3-5 on N GPUs (one node).   usage: scripts/gpu_multi_suite.sh <N> <tag> [workloads...]
N=$1; TAG=$2; shift 2
WL=${@:-cfg2 cfg3 cfg4 cfg5}
mkdir -p gpurun_out
P=29700
for w in $WL; do
  P=$((P+1))
  if [ "$w" = cfg2 ]; then EXTRA="--steps 100 --warmup 5 --no-e2e"; else EXTRA="--workload $w --steps 20 --warmup 3"; fi
  if [ "$N" = 1 ]; then timeout 900 python bench.py $EXTRA > gpurun_out/${TAG}_${w}_n$N.json 2> gpurun_out/${TAG}_${w}_n$N.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N $EXTRA > gpurun_out/${TAG}_${w}_n$N.json 2> gpurun_out/${TAG}_${w}_n$N.err; fi
  tail -1 gpurun_out/${TAG}_${w}_n$N.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$w', 'n', d['n_gpus'], 'value %.4g' % d['value'], 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d.get('gather_ok'), str(d['detail'])[:300])
except Exception as e: print('$w failed', e)
"
  tail -2 gpurun_out/${TAG}_${w}_n$N.err | cut -c1-300
done
