set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 600 gpurun_out/bench_c4.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_c4_reference.json 2>/dev/null; tail -c 400 gpurun_out/bench_c4_reference.json
python bench.py --workload C3 --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2>/dev/null
python bench.py --workload C2 --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2>/dev/null
python bench.py --workload C2x64 --steps 3 --warmup 1 > gpurun_out/bench_c2x64.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,smsp__inst_executed.sum,sm__cycles_elapsed.max,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,gpu__time_duration.sum,launch__registers_per_thread
timeout 400 ncu --metrics $M --clock-control none -k regex:"leg_|fft_" -c 6 --csv --page raw --log-file gpurun_out/ncu_c4_metrics.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_metrics.log 2>&1
ls -la gpurun_out | tail -15
