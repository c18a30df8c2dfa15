python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err; tail -c 400 gpurun_out/bench_c4_n2.json; tail -2 gpurun_out/bench_c4_n2.err
