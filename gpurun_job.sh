mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/test_gpu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/test_gpu_full.log
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_v4_n1.json 2> gpurun_out/bench_v4_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_v4_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v4.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_v4.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_v4_n1.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('metric', 'value', 'ms_per_step', 'gpu_launches', 'e2e', 'clocks')})
r = d['roofline']; print({k: r.get(k) for k in ('kernel', 'bound', 'achieved', 'peak', 'frac', 'kernel_ms')})
print(json.dumps(d.get('train_step'))[:200])
for a in d.get('also', []):
    print(a.get('metric', a.get('workload')), a.get('value'), a.get('ms_per_step'), (a.get('parity') or {}).get('max_rel'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_attention_v3.csv python bench.py --workload attention --steps 2 --warmup 1 --eager > /dev/null 2>&1
grep -c . gpurun_out/launches_attention_v3.csv
