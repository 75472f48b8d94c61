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


def test_dropout_generator_is_the_published_philox2x32_10():
    """The numpy restatement that the GPU test holds K2's inner-dropout mask to (tests/test_attention_dropout_gpu.py) against the known-answer
    vectors of Random123's philox2x32_10 (kat_vectors): the mask generator is the published Philox, not a look-alike."""
    import numpy as np
    from tests.test_attention_dropout_gpu import philox2x32_10
    for c0, c1, key, want in ((0x00000000, 0x00000000, 0x00000000, (0xff1dae59, 0x6cd10df2)),
                              (0xffffffff, 0xffffffff, 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
                              (0x243f6a88, 0x85a308d3, 0x13198a2e, (0xdd7ce038, 0xf62a4c12))):
        r0, r1 = philox2x32_10(np.array([c0], dtype=np.uint64), np.array([c1], dtype=np.uint64), key)
        assert (int(r0[0]), int(r1[0])) == want
