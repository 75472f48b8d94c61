"""Error budget of the two error-compensated tensor-core engines of K1a, emulated on the CPU (no GPU needed): the operand splits are exact
functions of the inputs, so the representation error of hi·hi + hi·lo + lo·hi can be checked against float64 without the kernel.  The GPU tests
(tests/test_gemm_tc_gpu.py) hold the kernels themselves to the same 1e-5 of the largest output."""
import torch


def _tf32(t):
    i = t.float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1fff).view(torch.float32)          # round to nearest on the 13 dropped mantissa bits (csrc/gemm_tc.cu tf32_round)


def _maxnorm_rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _three_products(xh, xl, wh, wl):
    xh, xl, wh, wl = (t.double() for t in (xh, xl, wh, wl))
    return xh @ wh.T + xh @ wl.T + xl @ wh.T


def _case(seed, M=256, K=2094, N=256, binary_cols=966):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    x[:, :binary_cols] = (torch.rand(M, binary_cols, generator=g) < 0.02).float()      # the multi-hot part of the profiles: exact in bf16
    w = torch.randn(N, K, generator=g) / K ** 0.5
    return x, w


def test_bf16_hi_lo_split_holds_the_fp32_parity_budget_at_k_2094():
    """B200REC_TC_BF16X3: hi = bf16(x), lo = bf16(x - hi) — 16 mantissa bits per operand, the lo·lo product dropped."""
    worst = 0.0
    for seed in range(3):
        for x, w in (_case(seed), _case(seed, binary_cols=0)):
            xh = x.bfloat16().float(); xl = (x - xh).bfloat16().float()
            wh = w.bfloat16().float(); wl = (w - wh).bfloat16().float()
            worst = max(worst, _maxnorm_rel(_three_products(xh, xl, wh, wl), x.double() @ w.double().T))
    assert 5e-7 < worst < 1e-5, worst                              # ~4e-6 on Gaussian data: inside the budget, not by a wide margin


def test_tf32_hi_lo_split_is_an_order_of_magnitude_tighter():
    """B200REC_TC_TF32X3: hi = tf32(x), lo = x - hi (truncated to TF32 by the MMA): ~21 mantissa bits per operand."""
    x, w = _case(7, binary_cols=0)
    xh = _tf32(x); xl = _tf32(x - xh)
    wh = _tf32(w); wl = _tf32(w - wh)
    err = _maxnorm_rel(_three_products(xh, xl, wh, wl), x.double() @ w.double().T)
    assert err < 5e-7, err
