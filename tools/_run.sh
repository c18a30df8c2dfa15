for v in "PIXSHT_BATCH_RA=4" "PIXSHT_BATCH_RA=2"; do env $v python bench.py --workload C2x64 --steps 3 --warmup 1 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), round(d['one_by_one_ms'],1))"; done
for v in "PIXSHT_R0A=6" "PIXSHT_R0A=8"; do for w in C4 C3; do env $v python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v $w', round(d['value'],2), [(k['kernel'], round(k['ms'],2)) for k in d['roofline']['kernels'] if 'anal<0' in k['kernel']])"; done; done
