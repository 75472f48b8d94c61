mkdir -p gpurun_out
B200REC_ATT_OVERLAP_PREPARE=1 timeout 600 python bench.py --workload attention > gpurun_out/bench_att_ov.json 2> gpurun_out/bench_att_ov.err; echo "rc=$?"
timeout 600 python bench.py --workload attention > gpurun_out/bench_att.json 2> gpurun_out/bench_att.err; echo "rc=$?"
python - <<'PY'
import json
for n in ('bench_att_ov', 'bench_att'):
    d = json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1])
    print(n, {k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d['parity']['max_rel'])
PY
