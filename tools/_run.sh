python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload C3 --steps 2 --warmup 2 > gpurun_out/bench_c3_n2.log 2>&1; tail -2 gpurun_out/bench_c3_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/bench_c4_n2.log 2>&1; tail -2 gpurun_out/bench_c4_n2.log
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
