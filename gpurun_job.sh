set -x
python -m pytest tests/test_models_gpu.py -m gpu -x -q > gpurun_out/t30.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/t30.log
python bench.py --workload k3hbm --no-cpu-baseline > gpurun_out/b30_k3hbm.json 2> gpurun_out/b30_k3hbm.err
python bench.py --workload graph --no-cpu-baseline --skip-hbm-regime > gpurun_out/b30_graph.json 2> gpurun_out/b30_graph.err
