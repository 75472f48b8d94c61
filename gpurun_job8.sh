set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/mp_peer_check.py > gpurun_out/mp_peer_check8.log 2>&1; echo "mp_peer_check8 rc=$?"
grep '^{' gpurun_out/mp_peer_check8.log | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 30 --warmup 5 --workload graph > gpurun_out/bench_graph_n8.json 2> gpurun_out/bench_graph_n8.err; echo "bench graph n8 rc=$?"
tail -c 900 gpurun_out/bench_graph_n8.json; tail -5 gpurun_out/bench_graph_n8.err
B200REC_PEER_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 30 --warmup 5 --workload graph > gpurun_out/bench_graph_n8_nooverlap.json 2> gpurun_out/bench_graph_n8_nooverlap.err; echo "bench graph n8 no-overlap rc=$?"
tail -c 500 gpurun_out/bench_graph_n8_nooverlap.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 30 --warmup 5 --workload graph --graph-scheme reduce > gpurun_out/bench_graph_n8_reduce.json 2> gpurun_out/bench_graph_n8_reduce.err; echo "bench graph n8 reduce rc=$?"
tail -c 500 gpurun_out/bench_graph_n8_reduce.json
