"""K7 (csrc/neg_sample.cu): batched negative sampling of the ranking datasets against the reference's rule
(src/neural_collaborative_filtering/datasets/base.py:57-78: p = r^w / sum(r^w), np.random.choice(ids, p=p)).  Random streams cannot be
bit-matched, so parity is defined GIVEN the uniforms: (1) the kernel's uniforms are the documented Philox2x32-10 construction, (2) fed with
the same uniforms, numpy's inverse-CDF rule (`cdf = cumsum(p); cdf /= cdf[-1]; cdf.searchsorted(u, side='right')`, what np.random.choice
does with p) picks the same list element for every sample, (3) empirical frequencies follow p."""
import numpy as np
import pandas as pd
import pytest
import torch

from tests.test_attention_dropout_gpu import philox2x32_10

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _frame(n, seed, max_len=300):
    rng = np.random.default_rng(seed)
    rows = []
    for k in range(n):
        m = 0 if k == 5 else int(rng.integers(1, max_len))
        ids = rng.choice(100_000, size=m, replace=False)
        r = rng.integers(0 if k % 9 == 0 else 1, 8, m) * 0.5          # some lists contain rating 0.0: probability 0 when w > 0
        if m and not r.any():
            r[0] = 1.0
        rows.append((int(rng.integers(0, 500)), int(rng.integers(0, 100_000)), ids, r))
    return pd.DataFrame(rows, columns=['userId', 'positive_movieId', 'negative_movieIds', 'negative_ratings'])


def _choice_given_u(ratings, w, u):
    """np.random.choice(p=...)'s rule (numpy/random/_generator.pyx / mtrand.pyx `choice`: cdf = p.cumsum(); cdf /= cdf[-1];
    idx = cdf.searchsorted(uniform, side='right')) with the reference's p (datasets/base.py:62-66)."""
    boosted = np.asarray(ratings, dtype=np.float64) ** w
    p = boosted / sum(boosted)
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side='right')), cdf


@pytest.mark.parametrize('w', [0.0, 1.0, 2.5])
def test_negative_sampling_given_the_uniforms(w):
    from deeprecommendation_b200.neural_collaborative_filtering.datasets.base import RankingDataset
    ds = RankingDataset(_frame(400, seed=3))
    ds.w = w
    sampler = ds.device_sampler(DEV, seed=0x1234_5678_9abc_def0)
    rows = np.random.default_rng(1).integers(0, 400, 1000)
    rows[:3] = 5                                                       # the sample without negatives
    sampler.offset = 2 ** 32 - 500                                     # the counter crosses 32 bits inside this batch
    off = sampler.offset
    users, pos, neg, lpos, u = sampler.sample(rows, return_uniforms=True)
    assert sampler.offset == off + 1000
    neg, lpos, u = neg.cpu().numpy(), lpos.cpu().numpy(), u.cpu().numpy()
    assert np.array_equal(users, ds.samples['userId'].to_numpy()[rows]) and np.array_equal(pos, ds.samples['positive_movieId'].to_numpy()[rows])
    ctr = off + np.arange(1000, dtype=np.uint64)
    r0, r1 = philox2x32_10(ctr & np.uint64(0xffffffff), (ctr >> np.uint64(32)) ^ np.uint64(0x12345678), 0x9abcdef0)
    want_u = ((r0 >> np.uint64(5)).astype(np.float64) * 67108864.0 + (r1 >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0
    assert np.array_equal(u, want_u)
    for b, row in enumerate(rows):
        ids, r = ds.samples['negative_movieIds'][row], ds.samples['negative_ratings'][row]
        if len(ids) == 0:
            assert neg[b] == -1 and lpos[b] == -1
            continue
        k, cdf = _choice_given_u(r, w, u[b])
        if lpos[b] != k:                                               # only a uniform within rounding of a CDF step may land one element off
            assert abs(lpos[b] - k) == 1 and np.min(np.abs(cdf - u[b])) < 1e-12, (b, lpos[b], k)
        assert neg[b] == ids[lpos[b]]
        assert w == 0.0 or r[lpos[b]] > 0                              # an element of probability 0 is never drawn


def test_negative_sampling_frequencies_and_stream_advance():
    from deeprecommendation_b200.neural_collaborative_filtering.datasets.base import RankingDataset
    frame = pd.DataFrame({'userId': [1], 'positive_movieId': [7], 'negative_movieIds': [np.arange(10, 15)],
                          'negative_ratings': [np.array([0.5, 1.0, 2.0, 0.0, 4.0])]})
    ds = RankingDataset(frame)
    ds.w = 2.0
    sampler = ds.device_sampler(DEV, seed=7)
    n = 200_000
    a = sampler.sample(np.zeros(n, dtype=np.int64))[2].cpu().numpy()
    b = sampler.sample(np.zeros(n, dtype=np.int64))[2].cpu().numpy()
    assert not np.array_equal(a, b)                                     # the second batch continues the stream
    p = np.array([0.25, 1.0, 4.0, 0.0, 16.0]) / 21.25
    freq = np.bincount(np.concatenate([a, b]) - 10, minlength=5) / (2 * n)
    assert freq[3] == 0.0 and np.max(np.abs(freq - p)) < 4e-3          # ~5 sigma of a 400k-sample binomial
