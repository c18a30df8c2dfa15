set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > gpurun_out/smoke_final.log; cat gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_c4_final.log 2>&1; tail -1 gpurun_out/bench_c4_final.log | cut -c1-400
python bench.py --workload C3 > gpurun_out/bench_c3_final.log 2>&1; tail -1 gpurun_out/bench_c3_final.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4_final.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__cycles_elapsed.max,lts__t_bytes.sum,smsp__warps_active.avg.per_cycle_active
ncu --metrics $M --clock-control none -k regex:"leg_|fft_" --launch-skip 6 --launch-count 6 --csv --log-file gpurun_out/metrics_c4_final.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_m.log 2>&1
