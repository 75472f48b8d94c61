set -x
python -m pytest tests/test_gemm_tc_gpu.py tests/test_models_gpu.py -m gpu -x -q > gpurun_out/t37.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t37.log
python bench.py --workload graph --no-cpu-baseline > gpurun_out/b37_graph.json 2> gpurun_out/b37_graph.err; echo "bench rc=$?"
tail -3 gpurun_out/b37_graph.err
