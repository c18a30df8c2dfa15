python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload C3 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c3_v5.log; python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_v5.log').read()); print(d['value'], d['stages'], d['roofline']['frac'], d['roofline_fft'], d['e2e'])"
python bench.py --workload C4 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c4_v5.log; python -c "
import json; d=json.loads(open('gpurun_out/bench_c4_v5.log').read()); print(d['value'], d['stages'], d['roofline']['frac'], d['roofline_fft'], d['e2e'])"
