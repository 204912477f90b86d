python tools/band_check.py 3 8 32 > gpurun_out/bc_sp.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg2_band4.json 2> gpurun_out/bench_cfg2_band4.err
