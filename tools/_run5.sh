python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg2_band8.json 2> gpurun_out/bench_cfg2_band8.err
python tools/one_fit.py cfg2 > gpurun_out/one_fit_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -c 2 -o gpurun_out/prof_tc_gemm_cfg2_c python tools/one_fit.py cfg2 > gpurun_out/ncu_full.log 2>&1
