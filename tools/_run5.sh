timeout 300 tools/tc_gemm_test perf > gpurun_out/tc_gemm_test6.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg2_band7.json 2> gpurun_out/bench_cfg2_band7.err
