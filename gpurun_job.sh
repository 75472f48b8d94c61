mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_models_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --workload basic > gpurun_out/b13_basic.json 2> gpurun_out/b13_basic.err; echo rc=$?; tail -2 gpurun_out/b13_basic.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/b13_basic.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('train_step'), d.get('cpu_baseline'))
P
