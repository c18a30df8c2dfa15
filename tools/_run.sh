python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "general or long_ring or golden or c1" 2>&1 | tail -3
python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_new.log 2>&1
tail -1 gpurun_out/bench_c4_new.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value']); print(d['roofline_fft']['achieved'], d['roofline_fft']['frac'])"
python bench.py --workload C3 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C3', d['value'], d['e2e']['value'], d['roofline_fft']['frac'])"
python bench.py --workload C2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C2', d['value'], d['e2e']['value'], d['roofline_fft']['frac'])"
