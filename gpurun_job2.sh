set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "bench n2 rc=$?"
tail -c 700 gpurun_out/bench_r02_n2.json; tail -4 gpurun_out/bench_r02_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 5 --warmup 3 --workload graph5 --graph5-scale 0.1 > gpurun_out/bench_graph5_s01_n2.json 2> gpurun_out/bench_graph5_s01_n2.err; echo "graph5 n2 rc=$?"
tail -c 900 gpurun_out/bench_graph5_s01_n2.json; tail -4 gpurun_out/bench_graph5_s01_n2.err
timeout 300 python -m pytest tests/test_edge_cases_gpu.py -x -q -m gpu 2>&1 | tail -3
