set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t32.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/t32.log
