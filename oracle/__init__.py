"""CPU oracle for the NeRAF acoustic-field hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a CPU restatement (torch CPU tensors / numpy, float64 where the
reference computes in float64) of the reference algorithm for the path named in
BASELINE.json: query encodings -> acoustic MLP -> spectral loss (+ gradients)
and Griffin-Lim ISTFT, plus the T60/EDT/C50 yard-sticks.  Every function cites
the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it, and only as the checker.  Nothing in
``neraf_b200/`` imports it; the product path has no CPU fallback.

Parity pinning: the reference ships no tests and no golden vectors (SURVEY.md
section 4 / 8c), so the oracle is pinned against outputs of the reference's OWN
classes run in the build container (``oracle/make_golden.py`` imports
``NeRAF.NeRAF_field.NeRAFAudioSoundField``, ``NeRAF.NeRAF_evaluator.STFTLoss``
and ``NeRAF.NeRAF_helper.measure_edt/measure_clarity`` from /root/reference
through ``oracle/refshim.py`` and ``torchaudio.transforms.GriffinLim``) and the
resulting vectors are committed under ``tests/golden/``.  The three pieces whose
arithmetic lives in packages that are NOT importable anywhere in this build
(nerfstudio ``NeRFEncoding`` / ``SceneBox.get_normalized_positions``,
tiny-cuda-nn ``SphericalHarmonics``, pyroomacoustics ``measure_rt60``) are
restated from their published algorithms: for those three functions parity is
UNPINNED (see DESIGN.md "Oracle").
"""
