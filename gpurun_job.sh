set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t43.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/t43.log
B200REC_ATT_BWD_SLICES=4 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -x -q -k "backward or gradients" > gpurun_out/t43_slices.log 2>&1; echo "pytest slices rc=$?"
tail -4 gpurun_out/t43_slices.log
python tools/att_bwd_bench.py > gpurun_out/att_bwd_bench_v2_s1.json 2> gpurun_out/att_bwd_bench_v2.err; cat gpurun_out/att_bwd_bench_v2_s1.json
B200REC_ATT_BWD_SLICES=4 python tools/att_bwd_bench.py > gpurun_out/att_bwd_bench_v2_s4.json 2>> gpurun_out/att_bwd_bench_v2.err; cat gpurun_out/att_bwd_bench_v2_s4.json
B200REC_ATT_BWD_SLICES=8 python tools/att_bwd_bench.py > gpurun_out/att_bwd_bench_v2_s8.json 2>> gpurun_out/att_bwd_bench_v2.err; cat gpurun_out/att_bwd_bench_v2_s8.json
python bench.py --workload attention --no-cpu-baseline --skip-hbm-regime > gpurun_out/b43_att.json 2> gpurun_out/b43_att.err
B200REC_ATT_COMPACT_STREAMING=1 python bench.py --workload attention --no-cpu-baseline --skip-hbm-regime --no-train-step > gpurun_out/b43_att_streaming.json 2> gpurun_out/b43_att_streaming.err
python bench.py --workload graph --no-cpu-baseline > gpurun_out/b43_graph.json 2> gpurun_out/b43_graph.err
