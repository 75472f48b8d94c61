mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 > gpurun_out/bench_v2_n8.json 2> gpurun_out/bench_v2_n8.err; echo "rc=$?"
tail -3 gpurun_out/bench_v2_n8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_v2_n8.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'e2e', 'scaling', 'host_binding')})
for a in d.get('also', []):
    print(a.get('metric', a.get('workload')), a.get('value'), a.get('ms_per_step'))
print(json.dumps(d.get('graph_scaling'))[:700])
PY
