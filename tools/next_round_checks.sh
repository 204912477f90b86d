#!/bin/bash
# First GPU call of the next round: run the two experiments written (but not run) at the end of r01 and re-measure the baseline.
#   make -C tools tc_gemm_skinny_test tc_gemm_tmema2_test tc_gemm_tmema_test tc_gemm_test tma_stream_bench      (here, before gpurun)
#   gpurun --timeout 600 -- 'bash tools/next_round_checks.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in tc_gemm_skinny_test tc_gemm_tmema2_test tc_gemm_tmema_test tc_gemm_test tma_stream_bench; do
  [ -x tools/$t ] || { echo "tools/$t is not built: run 'make -C tools $t' in the container before gpurun (no nvcc needed on the box, but the binary must travel)"; exit 1; }
done
timeout 120 tools/tma_stream_bench > gpurun_out/tma_stream_bench.log 2>&1; echo "tma_stream_bench rc=$?"; tail -40 gpurun_out/tma_stream_bench.log
for t in tc_gemm_skinny_test tc_gemm_tmema2_test tc_gemm_tmema_test; do
  timeout 120 tools/$t perf > gpurun_out/$t.log 2>&1; echo "$t rc=$?"; grep -E "OK|FAIL|perf|PASSED|FAILED|timeout|error" gpurun_out/$t.log | tail -12
done
timeout 150 tools/tc_gemm_test perf > gpurun_out/tc_gemm_test_r02.log 2>&1; echo "tc_gemm_test rc=$?"; grep -E "FAIL|PLANES|tc_gemm_test:" gpurun_out/tc_gemm_test_r02.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r02.log
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_cfg2_r02_start.json 2> gpurun_out/bench_cfg2_r02_start.err; echo "bench rc=$?"; cut -c1-220 gpurun_out/bench_cfg2_r02_start.json
# only meaningful when tc_gemm_skinny_test passed: the same bench with the turned-around Lanczos products, and the band route's parity tests under it
MVB_LB_SKINNY=1 timeout 300 python -m pytest tests -m gpu -x -q -k "trf or cfg2 or ill_conditioned or ragged" > gpurun_out/pytest_gpu_r02_skinny.log 2>&1; echo "pytest (skinny) rc=$?"; tail -3 gpurun_out/pytest_gpu_r02_skinny.log
MVB_LB_SKINNY=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_r02_skinny.json 2> gpurun_out/bench_cfg2_r02_skinny.err; echo "bench (skinny) rc=$?"; cut -c1-220 gpurun_out/bench_cfg2_r02_skinny.json
