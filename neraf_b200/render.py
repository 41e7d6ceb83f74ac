"""Batched render / eval driver: the audio half of ``ns-eval`` without its one-RIR-per-iteration loop.

The reference (/root/reference/NeRAF/NeRAF_pipeline.py:351-396) walks the eval set one RIR at a time: dataset item ->
``get_outputs_for_camera`` (ResNet3D + 60-100 field queries) -> ``get_image_metrics_and_images`` (two Griffin-Lim runs,
host copies, numpy metrics) -> ``np.save(eval_%05d.npy)``.  Here poses go through the field in chunks of hundreds of
RIRs (the grid feature computed once), Griffin-Lim and the acoustic metrics run once per chunk for all predictions and
targets, and the files keep the reference's name and layout: ``eval_00000.npy`` = float32 (C, F, T) log-STFT
(:371-377, what viz/loudness_maps.ipynb and viz/video.ipynb read).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


@torch.no_grad()
def render_poses(model, mic_poses, source_poses, rots, output_path: Optional[str] = None, chunk: int = 512,
                 waveforms: bool = False, init_phase: Optional[torch.Tensor] = None, first_index: int = 0) -> Dict:
    """N poses -> log-STFTs (N, C, F, T) on the host (and ``eval_%05d.npy`` files under ``output_path``), optionally the
    Griffin-Lim waveforms (N, C, hop*(T-1)) on the device (the loudness-map workload)."""
    mic = torch.as_tensor(mic_poses).reshape(-1, 3)
    src = torch.as_tensor(source_poses).reshape(-1, 3).expand(mic.shape[0], 3)
    rot = torch.as_tensor(rots).reshape(-1, 3).expand(mic.shape[0], 3)
    N = mic.shape[0]
    if output_path is not None:
        os.makedirs(output_path, exist_ok=True)
    g = model.grid_feature() if model.use_grid else None             # once for every pose (the reference: once per RIR)
    stfts, waves = [], []
    for lo in range(0, N, chunk):
        hi = min(lo + chunk, N)
        y = model.query_rirs(mic[lo:hi], src[lo:hi], rot[lo:hi], grid_feature=g)          # (n, T, C, F)
        if waveforms:
            ip = None if init_phase is None else init_phase[lo:hi]
            waves.append(model.istft_transform.render(y, ip))
        host = y.permute(0, 2, 3, 1).contiguous().cpu().numpy()                          # (n, C, F, T): raw_output.permute(1,2,0)
        stfts.append(host)
        if output_path is not None:
            for i in range(hi - lo):
                np.save(os.path.join(output_path, f"eval_{str(first_index + lo + i).zfill(5)}.npy"), host[i])
    out = {"stft": np.concatenate(stfts) if stfts else np.zeros((0, model.mic_ch, model.field.N_frequencies, model.max_len), np.float32)}
    if waveforms:
        out["wave"] = torch.cat(waves) if waves else None
    return out


@torch.no_grad()
def evaluate_rirs(model, items: Sequence[Dict], evaluator, chunk: int = 256, output_path: Optional[str] = None,
                  init_phase: Optional[torch.Tensor] = None) -> List[Dict[str, float]]:
    """The eval loop over dataset items ``{data (C,F,T) log-STFT, waveform (C,L_ff), mic_pose, source_pose, rot}``
    (``get_data_eval``, NeRAF_dataset.py:135-176): one metrics dict per RIR with the keys of
    ``evaluator.get_full_metrics`` (neraf_b200.evaluator), predictions optionally saved like ns-eval does."""
    results: List[Dict[str, float]] = []
    dev = model.device
    for lo in range(0, len(items), chunk):
        part = items[lo:lo + chunk]
        mic = torch.stack([torch.as_tensor(b["mic_pose"]) for b in part])
        src = torch.stack([torch.as_tensor(b["source_pose"]) for b in part])
        rot = torch.stack([torch.as_tensor(b["rot"]) for b in part])
        ip = None if init_phase is None else init_phase[lo:lo + len(part)]
        r = render_poses(model, mic, src, rot, output_path, chunk=len(part), waveforms=True, init_phase=ip, first_index=lo)
        gt_ff = torch.stack([torch.as_tensor(b["waveform"]) for b in part]).to(dev, torch.float32)
        log_gt = torch.stack([torch.as_tensor(b["data"]) for b in part]).to(dev, torch.float32)
        results.extend(evaluator.get_full_metrics_batch(gt_ff, r["wave"], log_gt))
    return results
