mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_edge_cases_gpu.py -x -q -m gpu 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mlp_tower -c 3 --csv python bench.py --workload attention --steps 1 --warmup 1 --eager 2>/dev/null | grep gpu__time | awk -F'","' '{print substr($5,1,50), $NF}' | tr -d '"'
timeout 600 python bench.py --workload attention 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('att', d.get('value'), d.get('ms_per_step'), d['parity']['max_rel'], d['roofline']['op_ms_per_batch'])"
