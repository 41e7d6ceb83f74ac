"""Oracle: acoustic yard-sticks T60 / EDT / C50 (test infrastructure only).

* C50  -- /root/reference/NeRAF/NeRAF_helper.py:104-107 (``measure_clarity``).
* EDT  -- NeRAF_helper.py:124-146 (``measure_edt``).
* T60  -- ``pyroomacoustics.experimental.measure_rt60`` (pyroomacoustics 0.7.3,
  README.md:48; call sites NeRAF_helper.py:58-59 decay_db=30 and :76 decay_db=10
  after ``torchaudio.functional.highpass_biquad(cutoff=200)`` :70-74).
  pyroomacoustics is not importable here: its published Schroeder-fit algorithm is
  restated -- parity UNPINNED for T60.
C50/EDT are pinned against the reference's own functions in tests/golden.
"""
from __future__ import annotations

import numpy as np


def measure_clarity(signal: np.ndarray, time: float = 50, fs: int = 44100) -> float:
    h2 = signal ** 2
    t = int((time / 1000) * fs + 1)
    return float(10 * np.log10(np.sum(h2[:t]) / np.sum(h2[t:])))


def _schroeder_db(h: np.ndarray) -> np.ndarray:
    """dtype-preserving like pyroomacoustics (``h = np.array(h); power = h ** 2``): a float32 response gives a float32
    running sum from the tail and a float32 curve."""
    power = np.array(h) ** 2
    energy = np.cumsum(power[::-1])[::-1]
    i_nz = np.max(np.where(energy > 0)[0])
    energy = energy[:i_nz]
    e_db = 10 * np.log10(energy)
    return e_db - e_db[0]


def measure_edt(h: np.ndarray, fs: float = 44100, decay_db: float = 10) -> float:
    h = np.array(h)
    power = h ** 2
    energy = np.cumsum(power[::-1])[::-1]
    if np.all(energy == 0):
        return float("nan")
    i_nz = np.max(np.where(energy > 0)[0])
    energy = energy[:i_nz]
    energy_db = 10 * np.log10(energy)
    energy_db -= energy_db[0]
    i_decay = np.min(np.where(-decay_db - energy_db > 0)[0])
    return float((60 / decay_db) * (i_decay / float(fs)))


def measure_rt60(h: np.ndarray, fs: float = 1, decay_db: float = 60) -> float:
    """pyroomacoustics.experimental.measure_rt60 restated (Schroeder integration, -5 dB .. -5-decay_db).  Raises
    ValueError (numpy's, on an empty ``np.where``) when the curve has no such span, like the original."""
    fs = float(fs)
    e_db = _schroeder_db(h)
    min_energy_db = -np.min(e_db)
    if min_energy_db - 5 < decay_db:
        decay_db = min_energy_db
    i_5db = np.min(np.where(e_db < -5)[0])
    t_5db = i_5db / fs
    i_decay = np.min(np.where(e_db < -5 - decay_db)[0])
    t_decay = i_decay / fs
    return float((60 / decay_db) * (t_decay - t_5db))


def highpass_biquad(x: np.ndarray, fs: float, cutoff: float = 200.0, q: float = 0.707) -> np.ndarray:
    """torchaudio.functional.highpass_biquad (RBJ cookbook high-pass), direct-form I, float64."""
    w0 = 2 * np.pi * cutoff / fs
    alpha = np.sin(w0) / 2.0 / q
    b0 = (1 + np.cos(w0)) / 2
    b1 = -1 - np.cos(w0)
    b2 = b0
    a0 = 1 + alpha
    a1 = -2 * np.cos(w0)
    a2 = 1 - alpha
    b0, b1, b2, a1, a2 = b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0
    y = np.zeros_like(x, dtype=np.float64)
    x1 = x2 = y1 = y2 = 0.0
    for n, xn in enumerate(np.asarray(x, dtype=np.float64)):
        yn = b0 * xn + b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2
        x2, x1 = x1, xn
        y2, y1 = y1, yn
        y[n] = yn
    return np.clip(y, -1.0, 1.0)          # torchaudio lfilter clamp=True default


def t60_soundspaces(h: np.ndarray, fs: float) -> float:
    """NeRAF_helper.py:48-64 (``compute_t60``, advanced=False): a failed fit reads -1."""
    try:
        return measure_rt60(h, fs=fs, decay_db=30)
    except (ValueError, IndexError):
        return -1.0


def t60_raf(h: np.ndarray, fs: float) -> float:
    """NeRAF_helper.py:48-77 (``compute_t60`` advanced=True -> ``measure_rt60_advance``).  The filtered signal is
    float32, as torchaudio returns it for a float32 waveform."""
    try:
        y = highpass_biquad(h, fs, 200.0)
        return measure_rt60(y.astype(np.asarray(h).dtype), fs, decay_db=10)
    except (ValueError, IndexError):
        return -1.0
