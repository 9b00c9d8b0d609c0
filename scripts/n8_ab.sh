N=${1:-8}
P=29800
run() { P=$((P+1)); env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 100 --warmup 5 --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$N $*', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['detail']['kernel_ms'].items()}, d['gather_ok'])"; }
shift
for cfg in "$@"; do run $cfg; done
