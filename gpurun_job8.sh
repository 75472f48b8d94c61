set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
B200REC_PEER_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 --workload graph > gpurun_out/bench_graph_n8_trace.json 2> gpurun_out/bench_graph_n8_trace.err; echo "bench graph n8 trace rc=$?"
tail -c 400 gpurun_out/bench_graph_n8_trace.json; tail -3 gpurun_out/bench_graph_n8_trace.err
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --steps 5 --warmup 3 --workload graph5 > gpurun_out/bench_graph5_n8.json 2> gpurun_out/bench_graph5_n8.err; echo "graph5 n8 rc=$?"
tail -c 1800 gpurun_out/bench_graph5_n8.json; tail -5 gpurun_out/bench_graph5_n8.err
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 8 --steps 5 --warmup 3 --workload allpairs_full > gpurun_out/bench_allpairs_full_n8.json 2> gpurun_out/bench_allpairs_full_n8.err; echo "allpairs_full n8 rc=$?"
tail -c 2500 gpurun_out/bench_allpairs_full_n8.json; tail -5 gpurun_out/bench_allpairs_full_n8.err
