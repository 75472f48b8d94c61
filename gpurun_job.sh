set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>&1; echo "ref rc=$?"
KREGEX='regex:gemm_tn|spmm_|mlp_tower|attention_pool|splitk|radix_|scan_|csr_|edge_|plan_|dinv|group_stats|pairhash|mask_targets|iota|_ids_'
python bench.py --workload graph --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_graph.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 300 --csv --log-file gpurun_out/launches_graph.csv python bench.py --workload graph --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_graph.log 2>&1
echo "ncu graph rc=$?"
python bench.py --workload attention --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_att.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 300 --csv --log-file gpurun_out/launches_attention.csv python bench.py --workload attention --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_att.log 2>&1
echo "ncu att rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spmm_chunk -s 2 -c 2 -o gpurun_out/prof_spmm python bench.py --workload graph --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_spmm.log 2>&1
echo "ncu full spmm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 4 -c 3 -o gpurun_out/prof_gemm_att python bench.py --workload attention --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_gemm.log 2>&1
echo "ncu full gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_pool -s 2 -c 2 -o gpurun_out/prof_attpool python bench.py --workload attention --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_att.log 2>&1
echo "ncu full attpool rc=$?"
ls -la gpurun_out
