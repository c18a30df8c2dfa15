python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not size" 2>&1 | tail -2
for w in C4 C3; do python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['value'], {k:v for k,v in d['e2e'].items() if 'ms' in k or k=='value' or 'maxabs' in k})"; done
