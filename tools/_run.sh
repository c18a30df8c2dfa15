python -m pytest tests -m gpu -x -q -k "not c4_size and not c3_size" 2>&1 | tail -3
for w in C3 C4; do python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'], d['e2e']['value'])"; done
python bench.py --workload C2 --steps 3 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_c2.log; python -c "
import json; d=json.loads(open('gpurun_out/bench_c2.log').read()); print('C2', d['value'], d['stages'], d['e2e']['value'], d['cpu_baseline'])"
