python -m pytest tests -m gpu -x -q -k "not c4_size and not c3_size" 2>&1 | tail -3
for w in C3 C4; do python bench.py --workload $w --steps 2 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'], d['roofline_fft']['frac'])"; done
