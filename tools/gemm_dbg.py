"""K1a wide form under the B200REC_TC_DBG switches (no MMA / no X loads / no W copies): which leg of the pipeline sets the pace.
usage: gemm_dbg.py MxKxN engine  (spawns one process per switch: the library reads the variable per call, results are garbage by design)"""
import os, sys, subprocess, json
if len(sys.argv) > 3:
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from deeprecommendation_b200 import ops
    M, K, N = (int(v) for v in sys.argv[1].split('x'))
    x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') / K ** 0.5; b = torch.randn(N, device='cuda')
    flush = torch.empty(256 * 1024 * 1024 // 4, device='cuda')
    ts = []
    for r in range(9):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.linear_raw(x, w, b, engine=sys.argv[2]); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    print(json.dumps({'dbg': int(os.environ.get('B200REC_TC_DBG', '0')), 'ms': round(sorted(ts)[len(ts) // 2], 4)}))
else:
    for dbg in (0, 1, 8, 16, 9, 17, 24, 25, 4):
        env = dict(os.environ, B200REC_TC_DBG=str(dbg))
        r = subprocess.run([sys.executable, __file__, sys.argv[1], sys.argv[2], 'child'], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-300:], flush=True)
