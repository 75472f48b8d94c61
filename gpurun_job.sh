set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_gpu.py tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_attention_r02_v1.csv python bench.py --workload attention --steps 2 --warmup 3 --no-cpu-baseline --no-train-step --eager > gpurun_out/ncu_att.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/launches_attention_r02_v1.csv | head -14
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_tc_kernel --launch-skip 8 --launch-count 2 -o gpurun_out/prof_gemm_tc_wide_r02 -f python bench.py --workload attention --steps 2 --warmup 3 --no-cpu-baseline --no-train-step --eager > gpurun_out/ncu_att_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_att_full.log
