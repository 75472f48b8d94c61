mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py tests/test_models_gpu.py -m gpu -q -x 2>&1 | tail -3
python tools/gemm_bench.py 2>&1 | grep -E "tf32x3|bf16"
python bench.py --no-cpu-baseline --workload attention > gpurun_out/b7.json 2> gpurun_out/b7.err; echo rc=$?; tail -2 gpurun_out/b7.err
