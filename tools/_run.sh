python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload C3 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c3_v6.log; python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_v6.log').read()); print(d['value'], d['stages'], d['e2e'])"
for K in 1 2 4 8; do PIXSHT_SPLITS=$K python bench.py --workload C4 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c4_v6_$K.log; python -c "
import json; d=json.loads(open('gpurun_out/bench_c4_v6_$K.log').read()); print($K, d['value'], d['stages'], d['e2e'])"; done
