C="python tools/gemm_bench.py"
$C > gpurun_out/plain_gb.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:node_gemm_kernel -s 2 -c 1 -f -o gpurun_out/prof_nodegemm_v1 $C > gpurun_out/ncu_ng.log 2>&1; echo "ncu rc=$?"
