"""CPU, world_size = 2 over gloo: the host-side data-parallel logic (batch / feature sharding, the
gradient all-reduce over the flat buffer, (sum, sumsq) statistics, feature-sharded Gram partials).
The local compute in these tests is the oracle (CPU); on the box the same helpers wrap the kernels."""
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alignq_b200.utils import sharding as S


from _dist_util import run2


def _grad_exchange(rank, world):
    torch.manual_seed(100 + rank)
    flat = torch.randn(1000)
    mine = flat.clone()
    S.allreduce_mean_(flat)
    return mine, flat


def test_flat_gradient_allreduce_is_the_mean_over_ranks():
    out = run2(_grad_exchange)
    mean = (out[0][0] + out[1][0]) / 2
    assert torch.allclose(out[0][1], mean) and torch.equal(out[0][1], out[1][1])


def _moments(rank, world):
    torch.manual_seed(7)
    w = torch.randn(10007) * 0.3 + 0.05                 # same tensor on every rank ...
    part = S.shard_batch(w, rank, world)                # ... each rank reduces only its shard
    m, s = S.combine_moments(part.double().sum(), (part.double() ** 2).sum(), torch.tensor(float(part.numel())))
    return float(m), float(s), float(w.mean()), float(w.std())


def test_sum_sumsq_allreduce_matches_single_device_stats():
    for m, s, m_ref, s_ref in run2(_moments):
        assert abs(m - m_ref) <= 1e-7 and abs(s / s_ref - 1) <= 1e-6


def _gram(rank, world):
    from oracle import alignq_oracle as O
    torch.manual_seed(3)
    B, F = 12, 1000                                      # ragged: 1000 is not a multiple of 2 * 32
    x = torch.randn(B, F, dtype=torch.float64)
    f0, f1 = S.feature_slice(F, rank, world)
    G = O.corr(x[:, f0:f1], x[:, f0:f1])                # local partial over ALL rows, own feature slice
    S.allreduce_gram_partial_(G, f1 - f0, F)
    return (f0, f1), G, O.corr(x, x)


def test_feature_sharded_gram_equals_single_device_gram():
    out = run2(_gram)
    assert out[0][0][1] == out[1][0][0] and out[0][0][0] == 0 and out[1][0][1] == 1000   # slices tile [0, F)
    for _, G, ref in out:
        assert torch.allclose(G, ref, rtol=1e-12, atol=1e-14)


def test_shard_batch_partitions_the_batch():
    x = torch.arange(131).view(131, 1)
    for world in (1, 2, 4, 8):
        parts = [S.shard_batch(x, r, world) for r in range(world)]
        assert torch.equal(torch.cat(parts), x)
        assert max(p.shape[0] for p in parts) - min(p.shape[0] for p in parts) <= 1
    for F in (32, 1000, 802816):
        for world in (1, 2, 8):
            sl = [S.feature_slice(F, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == F and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
