MVB_LB_PASS1_LOCAL=1 MVB_LB_REORTH2=0 timeout 900 python tools/band_check.py 64 > gpurun_out/bc_full_p1l_nore2.log 2>&1
timeout 900 python tools/band_check.py 64 > gpurun_out/bc_full_default.log 2>&1
