"""Oracle: the reference's training sample and its collated batch (test infrastructure only).

Restates ``SoundSpacesDataset.get_data`` (/root/reference/NeRAF/NeRAF_dataset.py:272-296) and
``RAFDataset.get_data`` from the point where the complex STFT exists (:113-130), plus torch's default collate as
the DataLoader of NeRAF_datamanager.py:84-91 applies it.  PINNED for the SoundSpaces path: tests/golden/
datafeed_SoundSpaces.npz holds batches produced by the reference's real class (oracle/make_golden_datafeed.py).
The RAF wav loader (librosa.load / resample, :96-109) is not restated: librosa is absent -- the cache is built from
STFT magnitudes, which is where both datasets meet.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np


def get_id_tmp(idx: int, max_len: int):
    """NeRAF_dataset.py:86-87 / :268-269."""
    return idx // max_len, idx % max_len


def target_column(mag: np.ndarray, t: int) -> np.ndarray:
    """mag: (C, F, T_file) STFT magnitudes of one RIR -> the (C, F) float32 target of time bin t.

    NeRAF_dataset.py:283-288: inside the recording log(mag[:, :, t] + 1e-3); past its end a constant column
    log(min(mag) + 1e-3)."""
    mag = np.asarray(mag, dtype=np.float32)
    if t < mag.shape[2]:
        return np.log(mag[:, :, t] + np.float32(1e-3)).astype(np.float32)
    col = np.ones(mag.shape[:2], dtype=np.float32) * mag.min()
    return np.log(col + np.float32(1e-3)).astype(np.float32)


def get_data(mags: Sequence[np.ndarray], mic: np.ndarray, src: np.ndarray, rot: np.ndarray, idx: int, max_len: int) -> Dict:
    rir, t = get_id_tmp(int(idx), max_len)
    return {"audio_idx": rir, "data": target_column(mags[rir], t), "time_query": t, "rot": rot[rir],
            "mic_pose": mic[rir], "source_pose": src[rir]}


def collate(samples: Sequence[Dict]) -> Dict[str, np.ndarray]:
    """torch.utils.data default_collate on these samples: python ints -> int64, float32 / float64 arrays stacked."""
    out = {}
    for k in samples[0]:
        v = [s[k] for s in samples]
        out[k] = np.asarray(v, dtype=np.int64) if isinstance(v[0], (int, np.integer)) else np.stack(v)
    return out


def batch(mags, mic, src, rot, indices, max_len) -> Dict[str, np.ndarray]:
    return collate([get_data(mags, mic, src, rot, i, max_len) for i in indices])


def full_cache(mags: Sequence[np.ndarray], max_len: int) -> np.ndarray:
    """(n_rirs * max_len, C * F) float32: every sample's target, row = dataset index."""
    rows = [target_column(m, t).reshape(-1) for m in mags for t in range(max_len)]
    return np.stack(rows)
