set -x
python -m pytest tests/test_edge_cases_gpu.py -m gpu -q > gpurun_out/t33.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/t33.log
