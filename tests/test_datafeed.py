"""Audio data feed: oracle and host logic on CPU, the gather kernel on the GPU.

The golden batches were produced by the reference's real ``SoundSpacesDataset`` + DataLoader
(oracle/make_golden_datafeed.py)."""
import os

import numpy as np
import pytest
import torch

from neraf_b200 import _lib
from neraf_b200.datafeed import EpochSampler, ResidentAudioFeed, target_columns
from oracle import datafeed as ofeed

from .util import cuda

KEYS = ("audio_idx", "data", "time_query", "rot", "mic_pose", "source_pose")


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "datafeed_SoundSpaces.npz"))
    mags = [z[f"mag{i}"] for i in range(len(z["lengths"]))]
    batches = [{k: z[f"b{b}:{k}"] for k in KEYS} for b in range(int(z["n_batches"]))]
    return z, mags, batches


def test_oracle_reproduces_the_reference_dataset_batches(golden_dir):
    z, mags, batches = _golden(golden_dir)
    idx, max_len = z["indices"], int(z["max_len"])
    for b, ref in enumerate(batches):
        got = ofeed.batch(mags, z["mic"], z["src"], z["rot"], idx[16 * b:16 * (b + 1)], max_len)
        for k in KEYS:
            assert got[k].dtype == ref[k].dtype and got[k].shape == ref[k].shape, k
            if k == "data":         # numpy's logf against torch's: at most one unit in the last place
                assert np.max(np.abs(got[k] - ref[k])) <= 2.4e-7 * np.max(np.abs(ref[k]))
            else:
                assert np.array_equal(got[k], ref[k]), k


def _equal_up_to_the_host_logf(a, ref):
    """Bit-equal where torch's logf of THIS host is the one that made the golden file; torch dispatches log to a
    different vector routine per ISA (AVX2 / AVX512 builds of SLEEF differ in the last bit on ~5 % of the inputs), so
    on another host the same arithmetic is allowed one unit in the last place."""
    return np.array_equal(a, ref) or np.max(np.abs(a - ref)) <= 1.2e-7 * np.max(np.abs(ref))


def test_cache_rows_equal_the_reference_samples_bit_for_bit(golden_dir):
    """Host side of the product: the cache is built with torch's own log, like the reference's samples."""
    z, mags, batches = _golden(golden_dir)
    max_len = int(z["max_len"])
    cache = torch.stack([target_columns(torch.from_numpy(m), max_len) for m in mags])       # (n, T, C, F)
    flat = cache.reshape(-1, cache.shape[2], cache.shape[3]).numpy()
    ref = np.concatenate([b["data"] for b in batches])
    assert _equal_up_to_the_host_logf(flat[z["indices"]], ref)
    # ... and bit-equal to the reference's own expression (NeRAF_dataset.py:283-288) evaluated by this host's torch
    for m, c in zip(mags, cache):
        t = min(m.shape[2], max_len)
        assert torch.equal(c[:t], torch.log(torch.from_numpy(m).float()[:, :, :t] + 1e-3).permute(2, 0, 1))
    assert np.max(np.abs(ofeed.full_cache(mags, max_len) - flat.reshape(flat.shape[0], -1))) < 3e-7 * np.abs(flat).max()


@pytest.mark.parametrize("drop_last", [False, True])
def test_two_ranks_see_what_one_rank_with_twice_the_batch_sees(drop_last):
    n, B = 1000, 48
    one = EpochSampler(n, 2 * B, seed=3, drop_last=drop_last)
    r0, r1 = (EpochSampler(n, B, seed=3, rank=r, world_size=2, drop_last=drop_last) for r in (0, 1))
    seen = []
    for _ in range(3 * (n // (2 * B) + 1)):
        e, lo, hi = one.next_range()
        whole = one.permutation(e)[lo:hi]
        parts = []
        for s in (r0, r1):
            es, l, h = s.next_range()
            assert es == e
            parts.append(s.permutation(es)[l:h])
        assert torch.equal(torch.cat(parts), whole)
        assert drop_last is False or len(whole) == 2 * B
        if e == 0:
            seen.append(whole)
    seen = torch.cat(seen)
    if drop_last:
        assert len(seen) == n // (2 * B) * 2 * B and len(set(seen.tolist())) == len(seen)
    else:
        assert sorted(seen.tolist()) == list(range(n))          # an epoch visits every (RIR, time bin) once


@pytest.mark.parametrize("n,B,world", [(10, 2, 2), (11, 2, 2), (9, 2, 2), (103, 8, 4), (7, 16, 3), (1000, 48, 8)])
def test_ranks_always_get_equal_shards(n, B, world):
    """The data-parallel loss normalises by n_local * world_size and every rank must reach its collectives: a partial
    last global batch is cut evenly (never an empty or a shorter shard on one rank), the < world_size left-over samples
    are skipped, and the shards stay disjoint rows of one permutation."""
    ranks = [EpochSampler(n, B, seed=1, rank=r, world_size=world) for r in range(world)]
    seen = []
    for _ in range(3 * (n // (B * world) + 2)):
        got = [s.next_range() for s in ranks]
        sizes = {hi - lo for _, lo, hi in got}
        assert len(sizes) == 1 and sizes.pop() >= 1 and len({e for e, _, _ in got}) == 1
        for r in range(1, world):
            assert got[r][1] == got[r - 1][2]                  # consecutive rows of the same global batch
        if got[0][0] == 0:
            seen.append(ranks[0].permutation(0)[got[0][1]:got[-1][2]])
    seen = torch.cat(seen)
    assert len(set(seen.tolist())) == len(seen) and n - len(seen) < world
    with pytest.raises(ValueError):
        EpochSampler(2, 4, world_size=3)


def test_feed_fails_loudly_without_a_device():
    with pytest.raises(_lib.NerafError):
        ResidentAudioFeed(torch.zeros(2, 4, 1, 5), torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 3), 4, 8)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gather_reproduces_the_reference_batches_bit_for_bit(golden_dir):
    dev = cuda()
    z, mags, batches = _golden(golden_dir)
    max_len = int(z["max_len"])
    feed = ResidentAudioFeed.from_magnitudes([torch.from_numpy(m) for m in mags], z["mic"], z["src"], z["rot"], max_len,
                                             16, dev)
    assert len(feed) == len(mags) * max_len
    for b, ref in enumerate(batches):
        got = feed.batch_from_indices(torch.from_numpy(z["indices"][16 * b:16 * (b + 1)]))
        for k in KEYS:
            assert got[k].dtype == torch.from_numpy(ref[k]).dtype, k
            if k == "data":
                assert _equal_up_to_the_host_logf(got[k].cpu().numpy(), ref[k]), k
                assert torch.equal(got[k].cpu(), feed.cache.cpu().reshape(-1, *got[k].shape[1:])[z["indices"][16 * b:16 * (b + 1)]])
            else:
                assert np.array_equal(got[k].cpu().numpy(), ref[k]), k
    feed.check()


@pytest.mark.gpu
def test_gather_at_training_size_against_the_oracle_and_into_static_buffers():
    """RAF-shaped cache (1, 513, 60) x 300 RIRs, B = 2048: every row equals the oracle's sample; filling
    pre-allocated (graph-static) buffers gives the same batch; an epoch visits every sample once."""
    dev = cuda()
    g = torch.Generator().manual_seed(11)
    n, T, C, F = 300, 60, 1, 513
    lengths = torch.randint(40, 75, (n,), generator=g).tolist()
    mags = [torch.rand(C, F, L, generator=g) * 2.0 for L in lengths]
    mic, src, rot = (torch.rand(n, 3, generator=g, dtype=torch.float64) for _ in range(3))
    feed = ResidentAudioFeed.from_magnitudes(mags, mic, src, rot, T, 2048, dev, seed=5)
    idx = torch.randint(0, n * T, (2048,), generator=g)
    got = feed.batch_from_indices(idx)
    ref = ofeed.batch([m.numpy() for m in mags], mic.numpy(), src.numpy(), rot.numpy(), idx.tolist(), T)
    for k in KEYS:
        a, r = got[k].cpu().numpy(), ref[k]
        if k == "data":
            assert np.max(np.abs(a - r)) <= 2.4e-7 * np.max(np.abs(r))        # torch logf (cache build) vs numpy logf
        else:
            assert np.array_equal(a, r), k
    static = {k: torch.empty_like(v) for k, v in got.items()}
    assert feed.batch_from_indices(idx, out=static) is static
    for k in KEYS:
        assert torch.equal(static[k], got[k]), k
    # all rows in order == the cache itself (size-independent property)
    everything = feed.batch_from_indices(torch.arange(n * T))
    assert torch.equal(everything["data"].reshape(n, T, C, F), feed.cache)
    assert torch.equal(everything["time_query"], torch.arange(n * T, device=dev) % T)
    # one epoch through next_train
    seen = []
    steps = -(-n * T // 2048)
    for s in range(steps):
        none, batch = feed.next_train(s)
        assert none is None
        seen.append(batch["audio_idx"] * T + batch["time_query"])
    seen = torch.cat(seen)
    assert seen.numel() == n * T and torch.equal(torch.sort(seen).values, torch.arange(n * T, device=dev))
    feed.check()


@pytest.mark.gpu
def test_two_rank_feeds_split_the_global_batch():
    """Data parallel: with one seed, rank r's batch is rows [r*B, (r+1)*B) of what a single feed with batch 2*B draws."""
    dev = cuda()
    n, T = 40, 12
    cache = torch.randn(n, T, 2, 9, device=dev)
    poses = [torch.rand(n, 3, dtype=torch.float64) for _ in range(3)]
    one = ResidentAudioFeed(cache, *poses, T, 64, seed=9)
    halves = [ResidentAudioFeed(cache, *poses, T, 32, seed=9, rank=r, world_size=2) for r in (0, 1)]
    for step in range(2 * (n * T // 64 + 1)):                     # across an epoch boundary, with a partial last batch
        whole = one.next_train(step)[1]
        parts = [f.next_train(step)[1] for f in halves]
        for k in KEYS:
            assert torch.equal(torch.cat([p[k] for p in parts]), whole[k]), (k, step)


@pytest.mark.gpu
def test_gather_edge_cases():
    dev = cuda()
    feed = ResidentAudioFeed(torch.arange(2 * 3 * 1 * 5, dtype=torch.float32, device=dev).reshape(2, 3, 1, 5),
                             torch.zeros(2, 3), torch.ones(2, 3), torch.full((2, 3), 2.0), 3, 4)
    empty = feed.batch_from_indices(torch.zeros(0, dtype=torch.int64))
    assert empty["data"].shape == (0, 1, 5)
    feed.check()
    bad = feed.batch_from_indices(torch.tensor([5, 6, -1, 0]))
    assert torch.equal(bad["data"][0, 0], torch.arange(25, 30, dtype=torch.float32, device=dev))
    assert torch.equal(bad["data"][1], bad["data"][3])          # out of range reads as index 0 ...
    with pytest.raises(IndexError):                              # ... and is reported
        feed.check()
    feed.check()                                                 # the flag was cleared
