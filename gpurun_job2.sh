mkdir -p gpurun_out
export NCCL_DEBUG=WARN
echo "== partition tests"; timeout 300 python -m pytest tests/test_models_gpu.py -m gpu -q -x -k "partition" 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_graph_check.py > gpurun_out/mp_check2.json 2> gpurun_out/mp_check2.err; echo "mp rc=$?"; tail -1 gpurun_out/mp_check2.json; tail -3 gpurun_out/mp_check2.err
for S in reduce gather; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload graph --graph-scheme $S > gpurun_out/bench_n2_$S.json 2> gpurun_out/bench_n2_$S.err; echo "bench2 $S rc=$?"; tail -2 gpurun_out/bench_n2_$S.err
done
