#!/bin/bash
# r02 final evidence: default bench, launch lists (one cfg2 fit; one batched configs[4] fold), ncu --set full of the new / changed kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_cfg2_r02_final.json 2> gpurun_out/bench_cfg2_r02_final.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_cfg2_r02_final.json | head -14
python tools/one_fit.py cfg2 > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_cfg2.csv python tools/one_fit.py cfg2 > gpurun_out/r02b_launches_cfg2.log 2>&1; echo "launch list cfg2 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_batched.csv python tools/one_batched.py > gpurun_out/r02b_launches_batched.log 2>&1; echo "launch list batched rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"chol_inv_kernel" -s 1 -c 1 -o gpurun_out/r02b_ncu_chol_inv python tools/one_batched.py > /dev/null 2>&1; echo "ncu chol rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel<128" -s 2 -c 1 -o gpurun_out/r02b_ncu_leverage python tools/one_batched.py > /dev/null 2>&1; echo "ncu leverage rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_ta_kernel<2" -s 40 -c 1 -o gpurun_out/r02b_ncu_skinny python tools/skinny_time.py > /dev/null 2>&1; echo "ncu skinny rc=$?"
for f in r02b_ncu_chol_inv r02b_ncu_leverage r02b_ncu_skinny; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out/r02b_* | head -20
