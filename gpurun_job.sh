set -x
B200REC_ATT_FUSED_MERGE=1 python -m pytest tests/test_models_gpu.py tests/test_edge_cases_gpu.py tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/t48_fused.log 2>&1; echo "pytest fused rc=$?"
tail -4 gpurun_out/t48_fused.log
python -m pytest tests/test_models_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q > gpurun_out/t48.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/t48.log
B200REC_ATT_FUSED_MERGE=1 python bench.py --workload attention --no-cpu-baseline --no-train-step > gpurun_out/b48_att_fused.json 2> gpurun_out/b48_att_fused.err
python bench.py --workload attention --no-cpu-baseline --no-train-step > gpurun_out/b48_att.json 2> gpurun_out/b48_att.err
