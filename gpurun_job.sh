set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --workload attention --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_att.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_pool -s 2 -c 1 -o gpurun_out/prof_attpool_v2 python bench.py --workload attention --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_att.log 2>&1
echo "ncu att rc=$?"
python bench.py --workload graph --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_graph.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:spmm_chunk -s 2 -c 1 -o gpurun_out/prof_spmm_v3 python bench.py --workload graph --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_spmm.log 2>&1
echo "ncu spmm rc=$?"
