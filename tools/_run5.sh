python tools/band_check.py 3 8 32 > gpurun_out/bc_cholblk.log 2>&1
python tools/band_stress.py > gpurun_out/band_stress.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_cfg2_band10.json 2> gpurun_out/bench_cfg2_band10.err
