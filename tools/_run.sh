python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not size" 2>&1 | tail -2
python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_new.log 2>&1
tail -1 gpurun_out/bench_c4_new.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value']); print([(k['kernel'], round(k['ms'],1), round(k['fp64_pipe_utilisation'],3)) for k in d['roofline']['kernels']])"
