for v in "PIXSHT_COPY_PRIO=-1 PIXSHT_HOST_PIECES=8" "PIXSHT_COPY_PRIO=0 PIXSHT_HOST_PIECES=8" "PIXSHT_COPY_PRIO=-1 PIXSHT_HOST_PIECES=3"; do
env $v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), round(d['e2e']['value'],1))"; done
