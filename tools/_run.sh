python -m pytest tests -m gpu -x -q -k "iqu_f64 or golden_spin2" 2>&1 | tail -2
python tools/loop_eff.py 3000 12 1.0 | grep "spin2"
python bench.py --workload C4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'], [ (k['kernel'], round(k['ms'],1)) for k in d['roofline']['kernels']])"
