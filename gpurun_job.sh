set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py -x -q -m gpu > gpurun_out/test_gemm.log 2>&1; echo "gemm tests rc=$?"
tail -8 gpurun_out/test_gemm.log
timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "gemm bench rc=$?"
grep -E "94|96" gpurun_out/gemm_bench.log | head -20
