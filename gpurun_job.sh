mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_dropout_gpu.py tests/test_models_gpu.py tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -12
