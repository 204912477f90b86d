python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg2_band9.json 2> gpurun_out/bench_cfg2_band9.err
