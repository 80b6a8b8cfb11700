run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --mode train --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        t=json.loads(l); print('$*', round(t['value']), round(t['ms_per_step'],2), round(t['e2e']['value']))
"; }
run X=1
run NCCL_MAX_CTAS=8
run NCCL_MAX_CTAS=2
