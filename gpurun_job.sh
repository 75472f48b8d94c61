set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_edge_cases_gpu.py tests/test_allpairs_gpu.py -x -q -m gpu > gpurun_out/test_models.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/test_models.log
