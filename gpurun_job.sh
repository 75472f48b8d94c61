mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_models_gpu.py tests/test_gemm_tc_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -4
echo "== bench attention"; timeout 300 python bench.py --workload attention --no-cpu-baseline > gpurun_out/b12_att.json 2> gpurun_out/b12_att.err; echo rc=$?; tail -2 gpurun_out/b12_att.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/b12_att.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d.get('e2e_resident'), d['roofline']['op_ms_per_step'])
P
