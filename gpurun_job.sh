mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== bench default"; timeout 900 python bench.py > gpurun_out/b8.json 2> gpurun_out/b8.err; echo rc=$?; tail -2 gpurun_out/b8.err
echo "== graph, no L2 keep"; B200REC_SPMM_L2_KEEP=0 timeout 300 python bench.py --workload graph --no-cpu-baseline > gpurun_out/b8_nokeep.json 2> gpurun_out/b8_nokeep.err; echo rc=$?
echo "== attention eager"; timeout 300 python bench.py --workload attention --no-cpu-baseline --eager > gpurun_out/b8_att_eager.json 2> gpurun_out/b8_att_eager.err; echo rc=$?
A="python bench.py --workload attention --no-cpu-baseline --steps 2 --warmup 3"
G="python bench.py --workload graph --no-cpu-baseline --steps 2 --warmup 3"
$A > gpurun_out/plain_att.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_attention_v3.csv $A > gpurun_out/ncu_att.log 2>&1; echo "ncu att rc=$?"
$G > gpurun_out/plain_graph.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_graph_v3.csv $G > gpurun_out/ncu_graph.log 2>&1; echo "ncu graph rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 2 -f -o gpurun_out/prof_gemm_tc_v3 $A > gpurun_out/ncu_full_gemm.log 2>&1; echo "ncu full gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_seg_kernel -s 4 -c 2 -f -o gpurun_out/prof_attseg_v3 $A > gpurun_out/ncu_full_att.log 2>&1; echo "ncu full att rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spmm_chunk_kernel -s 4 -c 2 -f -o gpurun_out/prof_spmm_v4 $G > gpurun_out/ncu_full_spmm.log 2>&1; echo "ncu full spmm rc=$?"
