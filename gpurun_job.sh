set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_dropout_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/test_drop.log 2>&1; echo "rc=$?"
tail -25 gpurun_out/test_drop.log
