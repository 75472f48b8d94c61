mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_models_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -3
python bench.py --no-cpu-baseline --skip-hbm-regime > gpurun_out/b6.json 2> gpurun_out/b6.err; echo rc=$?; tail -2 gpurun_out/b6.err
