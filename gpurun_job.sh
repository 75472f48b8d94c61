mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_models_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --workload graph --no-cpu-baseline > gpurun_out/b14_graph.json 2> gpurun_out/b14_graph.err; echo rc=$?; tail -2 gpurun_out/b14_graph.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/b14_graph.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['op_ms_per_step'])
P
