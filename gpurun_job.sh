mkdir -p gpurun_out
B200REC_GEMM_ENGINE=bf16x3 timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/test_gpu_bf16x3.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/test_gpu_bf16x3.log
timeout 600 python bench.py --workload attention --gemm bf16x3 > gpurun_out/bench_att_bf16x3.json 2> gpurun_out/bench_att_bf16x3.err; echo "rc=$?"
timeout 600 python bench.py --workload attention > gpurun_out/bench_att_tf32x3.json 2> gpurun_out/bench_att_tf32x3.err; echo "rc=$?"
python - <<'PY'
import json
for n in ('bf16x3', 'tf32x3'):
    d = json.loads(open(f'gpurun_out/bench_att_{n}.json').read().strip().splitlines()[-1])
    print(n, {k: d.get(k) for k in ('value', 'ms_per_step', 'parity', 'gpu_launches')})
    print(json.dumps(d['roofline'].get('op_ms_per_batch')), json.dumps(d['roofline'].get('other_kernels'))[:600])
PY
