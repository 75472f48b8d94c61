mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_attention_dropout_gpu.py tests/test_edge_cases_gpu.py tests/test_models_gpu.py -x -q -m gpu > gpurun_out/test_k2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/test_k2.log
timeout 600 python bench.py --workload attention > gpurun_out/bench_att.json 2> gpurun_out/bench_att.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_att.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d['parity']['max_rel'], d['e2e']['value'])
print(json.dumps(d['roofline'].get('op_ms_per_batch')))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv python bench.py --workload attention --steps 1 --warmup 1 --eager 2>/dev/null | grep gpu__time | awk -F'","' '{print substr($5,1,50), $NF}' | tr -d '"' | grep -i "wseg\|merge\|um_compact" | tail -6
timeout 600 python bench.py --workload k2hbm 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k2hbm', d.get('value'), d.get('ms_per_step'), d['roofline'].get('frac'))"
