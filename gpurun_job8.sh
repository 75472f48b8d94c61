set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
B200REC_PEER_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 --workload graph > gpurun_out/bench_graph_n8_trace2.json 2> gpurun_out/bench_graph_n8_trace2.err; echo "bench graph n8 trace rc=$?"
tail -c 400 gpurun_out/bench_graph_n8_trace2.json; tail -3 gpurun_out/bench_graph_n8_trace2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 30 --warmup 5 --workload graph > gpurun_out/bench_graph_n8_v2.json 2> gpurun_out/bench_graph_n8_v2.err; echo "bench graph n8 v2 rc=$?"
tail -c 500 gpurun_out/bench_graph_n8_v2.json
