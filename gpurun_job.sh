mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -x -q -m gpu -k "pipelined" 2>&1 | tail -12
timeout 900 python bench.py --workload attention > gpurun_out/bench_att.json 2> gpurun_out/bench_att.err; echo "rc=$?"; tail -2 gpurun_out/bench_att.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_att.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'e2e')}, d['parity']['max_rel'])
PY
