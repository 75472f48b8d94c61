set -x
python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q > gpurun_out/t34.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/t34.log
python bench.py --workload attention --no-cpu-baseline --skip-hbm-regime > gpurun_out/b34_att.json 2> gpurun_out/b34_att.err
python bench.py --workload basic --no-cpu-baseline --skip-hbm-regime > gpurun_out/b34_basic.json 2> gpurun_out/b34_basic.err
