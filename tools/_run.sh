python -m pytest tests -m gpu -x -q -k "golden or float32 or iqu_f64" 2>&1 | tail -2
for w in C3 C4; do python bench.py --workload $w --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'], d['roofline_fft']['frac'])"; done
