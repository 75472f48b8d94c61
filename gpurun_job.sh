set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload graph --steps 10 --warmup 3 --no-cpu-baseline --no-train-step > gpurun_out/bench_graph_n1.json 2> gpurun_out/bench_graph_n1.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/bench_graph_n1.json; tail -3 gpurun_out/bench_graph_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_graph_r02_v1.csv python bench.py --workload graph --steps 2 --warmup 3 --no-cpu-baseline --no-train-step --eager > gpurun_out/ncu_graph.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_graph_r02_v1.csv | tail -25
