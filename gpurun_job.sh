set -x
python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q > gpurun_out/t38.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/t38.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "backward and 300" > gpurun_out/t38_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/t38_memcheck.log
python tools/att_bwd_bench.py > gpurun_out/att_bwd_bench.json 2> gpurun_out/att_bwd_bench.err; cat gpurun_out/att_bwd_bench.json; tail -3 gpurun_out/att_bwd_bench.err
