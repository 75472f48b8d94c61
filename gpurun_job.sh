set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t43.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/t43.log
python tools/att_bwd_bench.py > gpurun_out/att_bwd_bench_v2.json 2> gpurun_out/att_bwd_bench_v2.err; cat gpurun_out/att_bwd_bench_v2.json
python bench.py --workload attention --no-cpu-baseline --skip-hbm-regime > gpurun_out/b43_att.json 2> gpurun_out/b43_att.err
B200REC_ATT_COMPACT_STREAMING=1 python bench.py --workload attention --no-cpu-baseline --skip-hbm-regime --no-train-step > gpurun_out/b43_att_streaming.json 2> gpurun_out/b43_att_streaming.err
