set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_allpairs_gpu.py -x -q -m gpu > gpurun_out/test_allpairs.log 2>&1; echo "allpairs tests rc=$?"
tail -6 gpurun_out/test_allpairs.log
timeout 600 python bench.py --workload allpairs_full --allpairs-users 125000 > gpurun_out/bench_allpairs_full_n1.json 2> gpurun_out/bench_allpairs_full_n1.err; echo "allpairs_full 1gpu rc=$?"
tail -c 3000 gpurun_out/bench_allpairs_full_n1.json; tail -3 gpurun_out/bench_allpairs_full_n1.err
