set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/mp_peer_check.py > gpurun_out/mp_peer_check2.log 2>&1; echo "mp_peer_check rc=$?"
grep '^{' gpurun_out/mp_peer_check2.log | tail -1; tail -5 gpurun_out/mp_peer_check2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 --workload graph > gpurun_out/bench_graph_n2.json 2> gpurun_out/bench_graph_n2.err; echo "bench graph n2 rc=$?"
tail -c 1500 gpurun_out/bench_graph_n2.json; tail -5 gpurun_out/bench_graph_n2.err
