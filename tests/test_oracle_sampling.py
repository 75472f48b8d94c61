"""The oracle's restatement of the ranking datasets' negative draw (oracle/restatement.py, datasets/base.py:57-78) pinned to the REAL
np.random.choice: under a seeded legacy generator, the element numpy draws equals the restated inverse-CDF rule applied to the uniform the
generator produces next.  The GPU test of K7 (tests/test_neg_sample_gpu.py) holds the kernel to the same rule."""
import numpy as np
import torch

from oracle import restatement as R


def test_choice_rule_is_numpys():
    rng = np.random.default_rng(0)
    for case in range(300):
        n = int(rng.integers(1, 60))
        r = rng.integers(0, 11, n) * 0.5
        if not r.any():
            r[0] = 3.0
        w = [0.0, 0.5, 1.0, 2.5][case % 4]
        p = R.negative_sampling_probs(r, w)
        assert abs(p.sum() - 1.0) < 1e-12 and (w == 0.0 or (p[r == 0] == 0).all())
        np.random.seed(case)
        drawn = int(np.random.choice(np.arange(n), p=p))
        np.random.seed(case)
        u = np.random.random_sample()
        assert drawn == R.choice_given_uniform(p, u)


def test_gpu_test_helper_is_the_oracle_rule():
    from tests.test_neg_sample_gpu import _choice_given_u
    rng = np.random.default_rng(1)
    for _ in range(200):
        r = rng.integers(1, 11, int(rng.integers(1, 40))) * 0.5
        u, w = float(rng.random()), float(rng.choice([0.0, 1.0, 2.5]))
        assert _choice_given_u(r, w, u)[0] == R.choice_given_uniform(R.negative_sampling_probs(r, w), u)


def test_bpr_loss_matches_the_mirrored_dataset_loss():
    from deeprecommendation_b200.neural_collaborative_filtering.datasets.base import BPR_loss
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(64, 1, generator=g), torch.randn(64, 1, generator=g)
    assert torch.allclose(BPR_loss(a, b), R.bpr_loss(a, b), rtol=1e-6, atol=1e-6)


def _ranking_fixture():
    from tests._golden import load
    d, _, _ = load('ranking_sampling')
    lists = [(d['neg_item'][a:b], d['neg_rating'][a:b]) for a, b in zip(d['neg_ptr'][:-1], d['neg_ptr'][1:])]
    return d, lists


def test_oracle_reproduces_the_reference_draws():
    """tests/golden/ranking_sampling.npz = the UNMODIFIED RankingDataset.__getitem__ (datasets/base.py:70-78) under a seeded legacy generator,
    with the uniform that generator yields recorded beside every draw (oracle/make_golden.py::golden_ranking)."""
    d, lists = _ranking_fixture()
    for w, tag in ((0.0, '0_0'), (1.0, '1_0'), (2.5, '2_5')):
        for item, u, want in zip(d['access'], d[f'uniform_w{tag}'], d[f'negative_w{tag}']):
            ids, r = lists[int(item)]
            assert ids[R.choice_given_uniform(R.negative_sampling_probs(r, w), float(u))] == want
    a, b = torch.from_numpy(d['bpr_pos']), torch.from_numpy(d['bpr_neg'])
    assert torch.allclose(R.bpr_loss(a, b), torch.from_numpy(d['bpr_loss']), rtol=1e-6)


def test_mirrored_ranking_dataset_draws_what_the_reference_draws():
    """the host path of the mirrored class (`RankingDataset.__getitem__`), same seeds -> same negatives"""
    import pandas as pd
    from deeprecommendation_b200.neural_collaborative_filtering.datasets.base import RankingDataset
    d, lists = _ranking_fixture()
    frame = pd.DataFrame({'userId': d['user'], 'positive_movieId': d['positive'], 'negative_movieIds': [ids for ids, _ in lists],
                          'negative_ratings': [r for _, r in lists]})
    ds = RankingDataset(frame)
    for w, tag in ((0.0, '0_0'), (1.0, '1_0'), (2.5, '2_5')):
        ds.w = w
        for t, (item, want) in enumerate(zip(d['access'], d[f'negative_w{tag}'])):
            np.random.seed(1000 + t)
            user, pos, neg = ds[int(item)]
            assert (user, pos, neg) == (d['user'][item], d['positive'][item], want)
    assert torch.allclose(ds.calculate_loss(torch.from_numpy(d['bpr_pos']), torch.from_numpy(d['bpr_neg'])), torch.from_numpy(d['bpr_loss']), rtol=1e-6)
