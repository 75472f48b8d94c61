set -x
python -m pytest tests/test_models_gpu.py tests/test_edge_cases_gpu.py tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/t36.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/t36.log
python bench.py --workload attention --no-cpu-baseline --skip-hbm-regime > gpurun_out/b36_att.json 2> gpurun_out/b36_att.err
B200REC_ATT_OVERLAP_PREPARE=1 python bench.py --workload attention --no-cpu-baseline --skip-hbm-regime > gpurun_out/b36_att_overlap.json 2> gpurun_out/b36_att_overlap.err
python bench.py --workload k2hbm --no-cpu-baseline > gpurun_out/b36_k2hbm.json 2> gpurun_out/b36_k2hbm.err
python bench.py --workload k2hbm --no-cpu-baseline > gpurun_out/b36_k2hbm_2.json 2> gpurun_out/b36_k2hbm_2.err
nvidia-smi --query-gpu=name,clocks.mem,clocks.max.mem,clocks.sm --format=csv > gpurun_out/b36_smi.txt
