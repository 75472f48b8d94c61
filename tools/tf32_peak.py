"""Dense TF32 and bf16 GEMM rates of this GPU through cuBLAS (torch.matmul), CUDA events, 8192^3 — the denominator of K1a's
tensor-pipe roofline in fp32-parity mode (MEASURED_PEAKS.json carries bf16 only).   python tools/tf32_peak.py > profiles/tf32_peak.json"""
import json

import torch


def rate(dtype, tf32, n=8192, reps=20):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device='cuda', dtype=dtype)
    b = torch.randn(n, n, device='cuda', dtype=dtype)
    for _ in range(5):
        a @ b
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        a @ b
    e.record()
    torch.cuda.synchronize()
    return 2.0 * n ** 3 * reps / (s.elapsed_time(e) * 1e-3) / 1e12


if __name__ == '__main__':
    out = {'tf32_tflops': round(rate(torch.float32, True), 1), 'bf16_tflops': round(rate(torch.bfloat16, False), 1),
           'fp32_simt_tflops': round(rate(torch.float32, False, n=4096, reps=10), 1),
           'what': 'cuBLAS via torch.matmul, 8192^3 (fp32 SIMT: 4096^3), 20 back-to-back launches between CUDA events, sustained',
           'gpu': torch.cuda.get_device_name(0)}
    print(json.dumps(out))
