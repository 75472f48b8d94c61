mkdir -p gpurun_out
python tools/gemm_one.py 9461x2094x256 'bf16x3!' || exit 1
ncu --set full --import-source on --clock-control none -k regex:"gemm_tc_kernel|tc_splitk_reduce" -s 2 -c 2 -f -o gpurun_out/ncu_gemm_bf16x3_v2 python tools/gemm_one.py 9461x2094x256 'bf16x3!' 4 > gpurun_out/ncu_g1.log 2>&1; tail -1 gpurun_out/ncu_g1.log
ncu --set full --import-source on --clock-control none -k regex:"gemm_tc_kernel" -s 1 -c 1 -f -o gpurun_out/ncu_gemm_tf32x3_v2 python tools/gemm_one.py 9461x2094x256 'tf32x3!' 4 > gpurun_out/ncu_g2.log 2>&1; tail -1 gpurun_out/ncu_g2.log
python tools/gemm_bench.py > gpurun_out/gemm_bench_v2.log 2>&1; tail -3 gpurun_out/gemm_bench_v2.log
