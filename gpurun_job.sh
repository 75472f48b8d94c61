set -x
for d in 8 9 10 11 15; do echo "== dbg $d"; B200REC_TC_DBG=$d python tools/gemm_bench.py 2>&1 | grep "9447x2094->256 \(tf32\|bf16\)"; done > gpurun_out/gb25_dbg.log 2>&1
