python tools/band_check.py 3 8 32 > gpurun_out/bc_chol.log 2>&1
MVB_LB_REORTH2=0 python tools/band_check.py 3 8 32 > gpurun_out/bc_chol_nore2.log 2>&1
MVB_LB_CHOL_TOL=0 python tools/band_check.py 3 8 32 > gpurun_out/bc_nochol.log 2>&1
