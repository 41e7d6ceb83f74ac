"""Eval-time metrics of rendered impulse responses: the evaluators of /root/reference/NeRAF/NeRAF_evaluator.py
(``RAFEvaluator`` :110-199, ``SoundSpacesEvaluator`` :201-262) with the waveforms kept on the device.

``get_full_metrics`` keeps the reference's signature and result keys for ONE RIR; ``get_full_metrics_batch`` measures
N RIRs in one pass (one ``neraf_acoustic_metrics`` launch for all 2*N*C waveforms) and returns one dict per RIR --
what the ns-eval loop (NeRAF_pipeline.py:351-396) accumulates one ``get_image_metrics_and_images`` call at a time.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from .metrics import acoustic_metrics


def _dev(x, device) -> torch.Tensor:
    return torch.as_tensor(x).to(device=device, dtype=torch.float32)


class _Evaluator:
    advanced = False
    t60_key = "audio_T60_mean_error"

    def __init__(self, fs: int, device=None):
        self.fs = fs
        self._device = None if device is None else torch.device(device)

    @property
    def device(self) -> torch.device:
        """Where the waveforms are measured: the constructor's device, else the current CUDA device -- resolved on first
        use, so that the model can build its evaluator on a host without a GPU (NeRAF_model.py:130,134)."""
        if self._device is None:
            self._device = torch.device("cuda", torch.cuda.current_device())
        return self._device

    # -- N RIRs at once ----------------------------------------------------------------------------
    def get_full_metrics_batch(self, wav_gt_ff, wav_pred_istft, log_gt=None) -> List[Dict[str, float]]:
        """wav_gt_ff: (N, C, L_ff) ground-truth waveforms from file; wav_pred_istft: (N, C, L) Griffin-Lim output of the
        prediction (L <= L_ff: zero-padded like NeRAF_evaluator.py:141-142); log_gt: (N, C, F, T) target log-STFT (RAF's
        round-trip error only)."""
        gt = _dev(wav_gt_ff, self.device)
        prd = _dev(wav_pred_istft, self.device)
        N, n_ch, L_ff = gt.shape
        prd = torch.nn.functional.pad(prd, (0, L_ff - prd.shape[-1]))
        m = acoustic_metrics(torch.stack([gt, prd]), self.fs, advanced=self.advanced)      # each (2, N, C)
        t60, edt, c50 = (m[k].cpu().numpy() for k in ("t60", "edt", "c50"))
        extra = self._extra(prd, log_gt)
        out = []
        for i in range(N):
            # NeRAF_evaluator.py:153-161 / :222-230: relative T60 error, 100 % for an RIR with a failed fit
            t60s = np.concatenate((t60[0, i], t60[1, i]))[None]
            with np.errstate(divide="ignore", invalid="ignore"):
                diff = np.abs(t60s[:, n_ch:] - t60s[:, :n_ch]) / np.abs(t60s[:, :n_ch])
            mask = np.any(t60s < -0.5, axis=1)
            diff = np.mean(diff, axis=1)
            diff[mask] = 1
            res = {self.t60_key: float(np.mean(diff) * 100), "audio_total_invalids_T60": float(np.sum(mask))}
            if extra is not None:
                res["audio_stft_error"] = float(extra[i])
            res["audio_EDT"] = float(np.mean(np.abs(edt[1, i] - edt[0, i]), axis=0))
            res["audio_C50"] = float(np.mean(np.abs(c50[1, i] - c50[0, i]), axis=0))
            out.append(res)
        return out

    def _extra(self, wav_prd: torch.Tensor, log_gt):
        return None

    # -- the reference's per-RIR call --------------------------------------------------------------
    def get_full_metrics(self, mag_prd, mag_gt, wav_gt_ff, wav_pred_istft, wav_gt_istft, log_prd, log_gt) -> Dict[str, float]:
        """Same arguments and keys as the reference; only wav_gt_ff, wav_pred_istft and log_gt are read (as there)."""
        lg = None if log_gt is None else torch.as_tensor(log_gt)[None]
        return self.get_full_metrics_batch(torch.as_tensor(wav_gt_ff)[None], torch.as_tensor(wav_pred_istft)[None], lg)[0]

    def get_stft_metrics(self, mag_prd, mag_gt):
        return {"audio_mag": torch.mean(torch.pow(mag_prd - mag_gt, 2)) * 2}


class SoundSpacesEvaluator(_Evaluator):
    """NeRAF_evaluator.py:201-262: T60 = measure_rt60(decay_db=30)."""

    def __init__(self, fs: int = 22050, device=None):
        super().__init__(fs, device)


class RAFEvaluator(_Evaluator):
    """NeRAF_evaluator.py:110-199: T60 after a 200 Hz high-pass (decay_db=10) and the STFT round-trip error."""
    advanced = True
    t60_key = "audio_T60"

    def __init__(self, fs: int = 48000, device=None):
        super().__init__(fs, device)
        if fs == 48000:
            self.n_fft, self.win_length, self.hop_len = 1024, 512, 256
        elif fs == 16000:
            self.n_fft, self.win_length, self.hop_len = 512, 256, 128
        else:
            raise ValueError("Sample rate not supported")          # NeRAF_evaluator.py:126

    def _extra(self, wav_prd: torch.Tensor, log_gt):
        """:144-149: back to an STFT from the (padded) predicted waveform, mean |log-magnitude difference| to the target."""
        if log_gt is None:
            return None
        lg = _dev(log_gt, self.device)
        N, C, L = wav_prd.shape
        win = torch.hann_window(self.win_length, device=self.device)
        spec = torch.stft(wav_prd.reshape(N * C, L), self.n_fft, self.hop_len, self.win_length, win, center=True,
                          pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
        log_prd = torch.log(spec.abs() + 1e-3).view(N, C, spec.shape[-2], spec.shape[-1])[..., :lg.shape[-1]]
        return (log_prd - lg).abs().mean(dim=(1, 2, 3)).cpu().numpy()

    def get_stft_metrics(self, mag_prd, mag_gt, gl=False):
        mag_loss = torch.mean(torch.pow(mag_prd - mag_gt, 2)) * 2
        spec_loss = torch.nn.functional.l1_loss(torch.log(1 + mag_prd), torch.log(1 + mag_gt)).item()     # SpectralLoss('mag', epsilon=1)
        return {"audio_mag": mag_loss, "audio_spectral_loss": spec_loss}
