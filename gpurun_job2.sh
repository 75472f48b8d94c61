mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_graph_check.py > gpurun_out/mp_check2.json 2> gpurun_out/mp_check2.err; echo "mp rc=$?"; tail -1 gpurun_out/mp_check2.json; tail -3 gpurun_out/mp_check2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; tail -2 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --impl reference > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; echo "ref2 rc=$?"; tail -2 gpurun_out/bench_n2_ref.err
