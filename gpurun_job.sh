set -x
python -m pytest tests/test_gemm_tc_gpu.py tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q > gpurun_out/t42.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/t42.log
B200REC_GEMM_ENGINE=tf32x3 python -m pytest tests -m gpu -q > gpurun_out/t42_tf32x3.log 2>&1; echo "pytest tf32x3-default rc=$?"
tail -15 gpurun_out/t42_tf32x3.log
python bench.py --workload basic --no-cpu-baseline > gpurun_out/b42_basic.json 2> gpurun_out/b42_basic.err
B200REC_MLP_TM4=0 python bench.py --workload basic --no-cpu-baseline > gpurun_out/b42_basic_tm8.json 2> gpurun_out/b42_basic_tm8.err
python bench.py --workload attention --no-cpu-baseline --skip-hbm-regime --no-train-step > gpurun_out/b42_att.json 2> gpurun_out/b42_att.err
