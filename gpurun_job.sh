set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_gpu.py -x -q -m gpu > gpurun_out/test_peer.log 2>&1; echo "peer tests rc=$?"
tail -5 gpurun_out/test_peer.log
timeout 600 python bench.py --workload graph5 --graph5-scale 0.004 --steps 5 > gpurun_out/bench_graph5_small_n1.json 2> gpurun_out/bench_graph5_small_n1.err; echo "graph5 small rc=$?"
tail -c 1500 gpurun_out/bench_graph5_small_n1.json; tail -5 gpurun_out/bench_graph5_small_n1.err
