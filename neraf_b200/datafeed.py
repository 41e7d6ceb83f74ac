"""GPU-resident audio data feed: the training batches of NeRAF without per-sample host work.

Mirrors, for ``mode='train'``, what ``SoundSpacesDataset`` / ``RAFDataset`` (/root/reference/NeRAF/NeRAF_dataset.py:
272-296, 89-132) plus the DataLoader of ``NeRAFDataManager`` (NeRAF_datamanager.py:80-119) hand to the model every
step -- the dict ``{audio_idx, data (B,C,F), time_query, rot, mic_pose, source_pose}`` -- but the ground-truth columns
live in HBM, laid out so that the dataset index is the cache row, and a batch is ONE gather kernel
(``neraf_gather_batch``).  The reference re-reads a (C,F,T) file (SoundSpaces) or re-STFTs a whole wav (RAF) for every
one of the 2048 columns of a step.

Sizes: RAF column 2 KB -> 123 KB per RIR; SoundSpaces 2 KB x 100 -> 206 KB per RIR; a 50 000-RIR scene is 6-10 GB of
180 GB.  Data parallel: every rank holds the whole cache and takes rows ``[rank*B, (rank+1)*B)`` of the global batch
of each step (same permutation on every rank: one seed), so N ranks see what one rank with batch N*B would.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib


def target_columns(mag: torch.Tensor, max_len: int) -> torch.Tensor:
    """(C, F, T_file) STFT magnitudes of one RIR -> (max_len, C, F) float32 training targets, one per time bin.

    NeRAF_dataset.py:283-288 (and :113-118): bins inside the recording are log(mag[:, :, t] + 1e-3); bins past its
    end are the constant column log(min(mag) + 1e-3)."""
    mag = mag.to(torch.float32)
    T = min(mag.shape[2], max_len)
    cols = torch.log(mag[:, :, :T] + 1e-3).permute(2, 0, 1)
    if T < max_len:
        pad = torch.log(torch.ones(mag.shape[0], mag.shape[1]) * mag.min() + 1e-3)
        cols = torch.cat([cols, pad.expand(max_len - T, -1, -1)], dim=0)
    return cols.contiguous()


class EpochSampler:
    """Index stream of DataLoader(shuffle=True, batch_size=B) (NeRAF_datamanager.py:84-91), shared by N ranks.

    Per epoch ONE permutation of the dataset (RandomSampler draws one torch.randperm), seeded from (seed, epoch) so
    every rank derives the same order; step s of the epoch is the global batch perm[s*N*B : (s+1)*N*B] and rank r
    owns its rows [r*B, (r+1)*B).  The last batch of an epoch is partial unless ``drop_last`` (DataLoader default);
    a new epoch starts when the permutation is used up (the reference restarts the loader on StopIteration, :111-116).
    Pure host logic: no device needed."""

    def __init__(self, n: int, batch_size: int, seed: int = 0, rank: int = 0, world_size: int = 1,
                 drop_last: bool = False):
        if n < world_size or batch_size <= 0 or not (0 <= rank < world_size):
            raise ValueError("EpochSampler: need n >= world_size > rank >= 0 and batch_size > 0")
        self.n, self.batch_size, self.seed, self.rank, self.world_size = n, batch_size, seed, rank, world_size
        self.drop_last = drop_last
        self.epoch, self.cursor = 0, 0

    def permutation(self, epoch: int) -> torch.Tensor:
        g = torch.Generator()
        g.manual_seed(self.seed * 1_000_003 + epoch)
        return torch.randperm(self.n, generator=g)

    def next_range(self) -> Tuple[int, int, int]:
        """(epoch, lo, hi): this rank's rows of the next global batch are permutation(epoch)[lo:hi].

        Every rank always gets the SAME number of rows: the data-parallel loss normalises by ``n_local * world_size``
        (loss.py) and a rank with an empty shard would never reach its collectives.  A partial last global batch of r
        rows is therefore cut into ``r // world_size`` rows per rank (the ``r % world_size`` < world_size samples left
        over are skipped for this epoch), and dropped altogether when that is zero."""
        gb = self.batch_size * self.world_size
        rem = self.n - self.cursor
        if rem <= 0 or (rem < gb and self.cursor > 0 and (self.drop_last or rem < self.world_size)):
            self.epoch, self.cursor = self.epoch + 1, 0
            rem = self.n
        per = self.batch_size if rem >= gb else rem // self.world_size
        lo = self.cursor + self.rank * per
        self.cursor += gb
        return self.epoch, lo, lo + per


class ResidentAudioFeed:
    """Training batches gathered on the device.  ``next_train(step)`` has the datamanager's signature."""

    def __init__(self, cache: torch.Tensor, mic_poses: torch.Tensor, source_poses: torch.Tensor, rots: torch.Tensor,
                 max_len: int, batch_size: int, seed: int = 0, rank: int = 0, world_size: int = 1,
                 drop_last: bool = False):
        if not cache.is_cuda:
            raise _lib.NerafError("ResidentAudioFeed: the cache must live on a CUDA device (there is no CPU path)")
        n_rirs = mic_poses.shape[0]
        if cache.dim() != 4 or cache.shape[0] != n_rirs or cache.shape[1] != max_len:
            raise ValueError("cache must be (n_rirs, max_len, C, F)")
        dev = cache.device
        self.cache = cache.to(torch.float32).contiguous()
        self.mic = mic_poses.to(device=dev, dtype=torch.float64).contiguous()
        self.src = source_poses.to(device=dev, dtype=torch.float64).contiguous()
        self.rot = rots.to(device=dev, dtype=torch.float64).contiguous()
        self.n_rirs, self.max_len, self.C, self.F = n_rirs, max_len, cache.shape[2], cache.shape[3]
        self.batch_size, self.seed, self.rank, self.world_size = batch_size, seed, rank, world_size
        self.drop_last = drop_last
        self.device = dev
        self.train_count = 0
        self.sampler = EpochSampler(n_rirs * max_len, batch_size, seed, rank, world_size, drop_last)
        self._epoch, self._perm = -1, None
        self._status = torch.zeros(1, dtype=torch.int32, device=dev)

    # -- construction ----------------------------------------------------------------------------
    @classmethod
    def from_magnitudes(cls, mags: Sequence[torch.Tensor], mic_poses, source_poses, rots, max_len: int,
                        batch_size: int, device, **kw) -> "ResidentAudioFeed":
        """mags[i]: (C, F, T_i) magnitudes (the SoundSpaces ``.npy`` files, or |Spectrogram(power=None)| of a RAF wav)."""
        cols = torch.stack([target_columns(torch.as_tensor(m), max_len) for m in mags])
        return cls(cols.to(device), torch.as_tensor(mic_poses), torch.as_tensor(source_poses), torch.as_tensor(rots),
                   max_len, batch_size, **kw)

    def __len__(self) -> int:                      # NeRAF_dataset.py:76-79 (train): one sample per (RIR, time bin)
        return self.n_rirs * self.max_len

    # -- sampling --------------------------------------------------------------------------------
    def _next_indices(self) -> torch.Tensor:
        epoch, lo, hi = self.sampler.next_range()
        if epoch != self._epoch:
            # the epoch's permutation is drawn ON the device (Philox, seeded like the host sampler: every rank gets the
            # same order) -- a host randperm of 10^5..10^7 indices plus its upload stalls the step for milliseconds
            g = torch.Generator(device=self.device)
            g.manual_seed(self.sampler.seed * 1_000_003 + epoch)
            self._epoch, self._perm = epoch, torch.randperm(len(self), generator=g, device=self.device)
        return self._perm[lo:hi]

    # -- gather ----------------------------------------------------------------------------------
    def batch_from_indices(self, idx: torch.Tensor, out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """The collated batch of dataset indices ``idx`` (int64, device).  ``out``: pre-allocated tensors to fill (the
        static buffers of a CUDA-graphed step), else new ones."""
        idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
        B, dev = idx.numel(), self.device
        if out is None:
            out = {"audio_idx": torch.empty(B, dtype=torch.int64, device=dev),
                   "data": torch.empty(B, self.C, self.F, dtype=torch.float32, device=dev),
                   "time_query": torch.empty(B, dtype=torch.int64, device=dev),
                   "rot": torch.empty(B, 3, dtype=torch.float64, device=dev),
                   "mic_pose": torch.empty(B, 3, dtype=torch.float64, device=dev),
                   "source_pose": torch.empty(B, 3, dtype=torch.float64, device=dev)}
        elif out["data"].shape[0] != B:
            raise ValueError(f"batch_from_indices: {B} indices for buffers of {out['data'].shape[0]} rows")
        aidx = out.get("audio_idx")
        _lib.check(_lib.lib().neraf_gather_batch(
            self.cache.data_ptr(), self.n_rirs, self.max_len, self.C * self.F, self.mic.data_ptr(), self.src.data_ptr(),
            self.rot.data_ptr(), idx.data_ptr(), B, out["data"].data_ptr(), out["time_query"].data_ptr(),
            _lib.ptr(aidx), out["mic_pose"].data_ptr(), out["source_pose"].data_ptr(), out["rot"].data_ptr(),
            self._status.data_ptr(), _lib.stream_ptr(dev)))
        return out

    def check(self) -> None:
        """Raises if any index handed to the gather so far was outside the dataset (device-side flag, one sync)."""
        if int(self._status.item()) != 0:
            self._status.zero_()
            raise IndexError("ResidentAudioFeed: a sample index was outside [0, len)")

    def next_train(self, step: int, out: Optional[Dict[str, torch.Tensor]] = None) -> Tuple[None, Dict[str, torch.Tensor]]:
        """NeRAF_datamanager.py:107-121: returns (None, batch)."""
        self.train_count += 1
        return None, self.batch_from_indices(self._next_indices(), out)
