set -x
python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_graph_build_gpu.py -m gpu -x -q > gpurun_out/t27.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/t27.log
python bench.py --workload k3hbm --no-cpu-baseline > gpurun_out/b27_k3hbm.json 2> gpurun_out/b27_k3hbm.err
python bench.py --workload graph --no-cpu-baseline --skip-hbm-regime > gpurun_out/b27_graph.json 2> gpurun_out/b27_graph.err
