set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
for cfg in "4 4" "4 2" "8 4" "2 2"; do set -- $cfg; echo "R0=$1 R2=$2"; PIXSHT_R0=$1 PIXSHT_R2=$2 python bench.py --workload C3 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'], d['roofline']['frac'])"; done > gpurun_out/rsweep2.log 2>&1
cat gpurun_out/rsweep2.log
python bench.py --workload C4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c4_v4.log 2>&1; tail -1 gpurun_out/bench_c4_v4.log
