set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/test_gpu_all.log 2>&1; echo "gpu tests rc=$?"
tail -6 gpurun_out/test_gpu_all.log
timeout 900 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r02_n1.json; tail -3 gpurun_out/bench_r02_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
