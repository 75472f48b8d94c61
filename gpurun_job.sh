set -x
python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -x -q > gpurun_out/t18.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t18.log
python bench.py --workload k2hbm --no-cpu-baseline > gpurun_out/b18_k2hbm.json 2> gpurun_out/b18_k2hbm.err
python bench.py --workload attention --no-cpu-baseline --skip-hbm-regime > gpurun_out/b18_att.json 2> gpurun_out/b18_att.err
