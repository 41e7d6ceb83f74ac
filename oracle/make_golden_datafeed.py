"""Generate tests/golden/datafeed_SoundSpaces.npz with the REFERENCE's real dataset class (build container only).

    python -m oracle.make_golden_datafeed

Writes a handful of synthetic (2, 257, T_i) magnitude files (some shorter than max_len) to a temp directory, runs
``SoundSpacesDataset(mode='train')`` from /root/reference/NeRAF/NeRAF_dataset.py through a torch DataLoader exactly as
NeRAF_datamanager.py:84-91 does (batch_size, default collate) on a fixed index list, and stores inputs + batches.
Test infrastructure only.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim     # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    _, SoundSpacesDataset = refshim.load_datasets()
    rng = np.random.default_rng(5)
    max_len, C, F = 12, 2, 257
    lengths = [12, 9, 15, 12, 5, 11]                      # shorter, equal and longer than max_len
    n = len(lengths)
    names = [f"{90 * (i % 4)}/{i}_{i + 1}" for i in range(n)]
    mags = [np.abs(rng.standard_normal((C, F, T))).astype(np.float32) * np.exp(-np.arange(T) / 4.0).astype(np.float32)
            for T in lengths]
    mic = rng.uniform(-3, 3, (n, 3))
    src = rng.uniform(-3, 3, (n, 3))
    rot = (np.stack([np.cos(np.arange(n) * np.pi / 2), np.zeros(n), np.sin(np.arange(n) * np.pi / 2)], 1) + 1.0) / 2.0
    with tempfile.TemporaryDirectory() as tmp:
        for name, m in zip(names, mags):
            os.makedirs(os.path.dirname(os.path.join(tmp, name)), exist_ok=True)
            np.save(os.path.join(tmp, name + ".npy"), m)
        outputs = types.SimpleNamespace(audios_filenames=names, microphone_poses=torch.from_numpy(mic),
                                        source_poses=torch.from_numpy(src), microphone_rotations=torch.from_numpy(rot),
                                        scene_box=None)
        ds = SoundSpacesDataset(outputs, mode="train", max_len=max_len, mag_path=tmp, wav_path=tmp)
        assert len(ds) == n * max_len
        indices = rng.permutation(len(ds))[:40].tolist() + [0, len(ds) - 1, 4 * max_len + 5, 4 * max_len + 11]
        loader = torch.utils.data.DataLoader(torch.utils.data.Subset(ds, indices), batch_size=16, shuffle=False)
        batches = list(loader)
    out = {"max_len": np.array(max_len), "lengths": np.array(lengths), "mic": mic, "src": src, "rot": rot,
           "indices": np.array(indices, dtype=np.int64), "n_batches": np.array(len(batches))}
    for i, m in enumerate(mags):
        out[f"mag{i}"] = m
    for b, batch in enumerate(batches):
        for k, v in batch.items():
            out[f"b{b}:{k}"] = v.numpy()
            if b == 0:
                print(k, v.dtype, tuple(v.shape))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "datafeed_SoundSpaces.npz"), **out)
    print("datafeed golden:", len(batches), "batches,", len(indices), "samples")


if __name__ == "__main__":
    main()
