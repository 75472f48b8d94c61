set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t49.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/t49.log
python bench.py --workload attention --no-cpu-baseline --skip-hbm-regime --no-train-step > gpurun_out/b49_att.json 2> gpurun_out/b49_att.err
