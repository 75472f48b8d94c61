set -x
python tools/train_profile.py > gpurun_out/train_profile_v2.txt 2>&1; grep "ms/step" gpurun_out/train_profile_v2.txt
B200REC_SPLITK_ANY_N=0 python tools/train_profile.py > gpurun_out/train_profile_v2_n4only.txt 2>&1; grep "ms/step" gpurun_out/train_profile_v2_n4only.txt
python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -x -q -k "backward or gradients" > gpurun_out/t44.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t44.log
