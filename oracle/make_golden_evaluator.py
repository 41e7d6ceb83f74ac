"""Generate tests/golden/evaluator_{RAF,SoundSpaces}.npz with the REFERENCE's real evaluator classes (build container only).

    python -m oracle.make_golden_evaluator

``RAFEvaluator`` / ``SoundSpacesEvaluator`` (NeRAF_evaluator.py:110-262) run unmodified; the one function they reach
that is not importable here, ``pyroomacoustics.experimental.measure_rt60``, is supplied by its restatement in
oracle/metrics.py (so T60 stays "parity unpinned", everything around it -- padding, error formulas, EDT, C50, the RAF
STFT round trip -- is the reference's own code).  Test infrastructure only.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neraf_b200 import synthetic as syn          # noqa: E402
from oracle import metrics as omet                # noqa: E402
from oracle import refshim                        # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def case(shape, Evaluator, n, seed):
    rng = np.random.default_rng(seed)
    L = shape.hop * (shape.T - 1)
    L_ff = shape.hop * shape.T if shape.C == 2 else int(0.32 * shape.fs)          # max_len_time of the datasets
    t60 = rng.uniform(0.1, 0.6, (n, 1, 1))
    env = np.exp(-6.91 * np.arange(L_ff)[None, None] / (t60 * shape.fs))
    gt_ff = (rng.standard_normal((n, shape.C, L_ff)) * env * 0.5).astype(np.float32)
    prd = (gt_ff[..., :L] * rng.uniform(0.7, 1.3, (n, shape.C, 1)) +
           0.02 * rng.standard_normal((n, shape.C, L)) * env[..., :L]).astype(np.float32)
    prd[1, 0] = 0.0                                                                  # a silent prediction: invalid T60
    log_gt = (rng.standard_normal((n, shape.C, shape.F, shape.T)) - 3.0).astype(np.float32)
    ev = Evaluator(fs=shape.fs)
    keys, rows = None, []
    for i in range(n):
        res = ev.get_full_metrics(None, None, gt_ff[i], prd[i], prd[i], None, log_gt[i])
        keys = keys or list(res)
        rows.append([res[k] for k in keys])
        print(shape.name, i, res)
    np.savez_compressed(os.path.join(OUT, f"evaluator_{shape.name}.npz"), gt_ff=gt_ff, prd=prd, log_gt=log_gt,
                        keys=np.array(keys), rows=np.array(rows, dtype=np.float64))


def main():
    refshim.install()
    import pyroomacoustics
    pyroomacoustics.experimental.measure_rt60 = lambda h, fs=1, decay_db=60, plot=False, **kw: omet.measure_rt60(h, fs, decay_db)
    from NeRAF.NeRAF_evaluator import RAFEvaluator, SoundSpacesEvaluator
    import NeRAF.NeRAF_helper as helper
    helper.pyroomacoustics = pyroomacoustics
    case(syn.RAF, RAFEvaluator, 4, 0)
    case(syn.SOUNDSPACES, SoundSpacesEvaluator, 4, 1)


if __name__ == "__main__":
    main()
