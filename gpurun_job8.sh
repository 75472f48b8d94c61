mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/mp_graph_check.py > gpurun_out/mp_check8.json 2> gpurun_out/mp_check8.err; echo "mp rc=$?"; tail -1 gpurun_out/mp_check8.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench8 rc=$?"; tail -2 gpurun_out/bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 --workload graph > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench4 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 --workload graph --graph-scheme gather > gpurun_out/bench_n8_gather.json 2> gpurun_out/bench_n8_gather.err; echo "bench8 gather rc=$?"
