python tools/train_profile.py > gpurun_out/train_profile.txt 2>&1; echo rc=$?
