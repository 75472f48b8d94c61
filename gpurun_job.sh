mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x 2>&1 | tail -3
for d in 0 7; do echo "== DBG $d"; B200REC_TC_DBG=$d python tools/gemm_bench.py 2>&1 | grep -E "tf32x3|bf16"; done
