#!/bin/bash
# r02 evidence: GPU tests, bench, launch list of one cfg2 fit, ncu --set full of the Gram / Z / sweep / score kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r02_full.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg2_r02.json 2> gpurun_out/bench_cfg2_r02.err; echo "bench rc=$?"
python tools/one_fit.py cfg2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg2.csv python tools/one_fit.py cfg2 > gpurun_out/r02_launches_cfg2.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"toeplitz_gram_kernel" -c 1 -o gpurun_out/r02_ncu_gram python tools/one_fit.py cfg2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_ta_kernel<0>|sweep_chain_tmem" -c 5 -o gpurun_out/r02_ncu_zr_sweep python tools/one_fit.py cfg2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:score_vec -c 1 -o gpurun_out/r02_ncu_score python tools/score_probe.py 1 > /dev/null 2>&1
for f in r02_ncu_gram r02_ncu_zr_sweep r02_ncu_score; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out/r02_* | head
