python tools/one_fit.py cfg2 > gpurun_out/one_fit_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg2_fold_r01c.csv python tools/one_fit.py cfg2 > gpurun_out/ncu_one_fit.log 2>&1
python tools/one_fit.py cfg2 > gpurun_out/one_fit_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -c 2 -o gpurun_out/prof_tc_gemm_cfg2_c python tools/one_fit.py cfg2 > gpurun_out/ncu_full.log 2>&1
echo done
