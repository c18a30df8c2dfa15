python -m pytest tests -m gpu -x -q -k "not c4_size and not c3_size" 2>&1 | tail -2
for f in 16 1; do for w in C3 C4; do PIXSHT_FFT_FUSE=$f python bench.py --workload $w --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fuse', $f, d['value'], d['stages'], round(d['roofline_fft']['frac'],3))"; done; done
python bench.py --workload C2 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', d['value'], d['stages'], d['e2e']['value'])"
