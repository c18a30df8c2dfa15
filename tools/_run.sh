python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"fft_" --launch-skip 2 --launch-count 2 -f -o gpurun_out/prof_fft_c4 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fft.log 2>&1
ncu -i gpurun_out/prof_fft_c4.ncu-rep --page raw --csv > gpurun_out/prof_fft_c4_raw.csv
ncu -i gpurun_out/prof_fft_c4.ncu-rep --page source --csv --kernel-name regex:fft_phase2map > gpurun_out/prof_fft_c4_src.csv
