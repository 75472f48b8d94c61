mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_gemm_tc_gpu.py -m gpu -q -x 2>&1 | tail -4
echo "== bench attention"; timeout 300 python bench.py --workload attention --no-cpu-baseline > gpurun_out/b11_att.json 2> gpurun_out/b11_att.err; echo rc=$?; tail -2 gpurun_out/b11_att.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/b11_att.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['op_ms_per_step'])
P
A="python bench.py --workload attention --no-cpu-baseline --steps 2 --warmup 3"
$A > gpurun_out/plain_att.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_attention_v5.csv $A > gpurun_out/ncu_att.log 2>&1; echo "ncu att rc=$?"
