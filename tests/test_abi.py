"""CPU tests of the C-ABI boundary: the library builds, loads and exports exactly the symbols the header
declares; argument validation that needs no GPU; the Griffin-Lim core checked lane-by-lane on the host."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "neraf_b200.h")).read()
    return sorted(set(re.findall(r"NERAF_API\s+[\w\s\*]+?\b(neraf_\w+)\s*\(", text)))



def test_library_exports_every_declared_symbol(built):
    from neraf_b200 import _lib
    names = _header_symbols()
    assert len(names) >= 16
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (neraf_\w+)", out)))
    assert exported == names
    lib = _lib.lib()
    assert lib.neraf_version() == 9


def test_ctypes_mirrors_have_the_layout_of_the_header(built):
    """Every struct the header declares has a ctypes mirror of the same size (a field added on one side only would
    silently shift everything behind it)."""
    from neraf_b200 import _lib
    lib = _lib.lib()
    text = open(os.path.join(ROOT, "include", "neraf_b200.h")).read()
    declared = sorted(set(re.findall(r"^}\s*(neraf_\w+);", text, flags=re.M)))
    assert declared == sorted(_lib.STRUCTS), "a public struct has no ctypes mirror"
    for name, cls in _lib.STRUCTS.items():
        assert lib.neraf_abi_sizeof(name.encode()) == C.sizeof(cls), name
    assert lib.neraf_abi_sizeof(b"no_such_struct") == 0


def test_library_is_sm100a_with_tcgen05_and_tma(built):
    from neraf_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_size_queries_and_validation_without_gpu(built):
    from neraf_b200 import _lib
    lib = _lib.lib()
    dims = _lib.make_dims(1024, 163, [5096, 2048, 1024, 1024, 512], 1, 513)
    pack, ws = C.c_size_t(), C.c_size_t()
    assert lib.neraf_field_sizes(C.byref(dims), _lib.PREC_BF16, 2048, C.byref(pack), C.byref(ws)) == 0
    n_w = 163 * 5096 + 5096 * 2048 + 2048 * 1024 + 1024 * 1024 + 1024 * 512 + 512 * 513
    assert 2 * n_w <= pack.value < 2 * n_w * 1.05                  # one bf16 copy per weight matrix (+ padding)
    assert ws.value > 2048 * 9704 * 2 * 2                          # bf16 activations + their gradients, no transposes
    assert lib.neraf_field_sizes(C.byref(dims), _lib.PREC_FP32, 2048, C.byref(pack), C.byref(ws)) == 0
    assert pack.value == 0
    assert lib.neraf_field_sizes(C.byref(dims), 7, 2048, C.byref(pack), C.byref(ws)) == 1
    assert b"precision" in lib.neraf_last_error()
    bad = _lib.make_dims(1024, 163, [], 1, 513)
    assert lib.neraf_field_sizes(C.byref(bad), 0, 1, C.byref(pack), C.byref(ws)) == 1
    p = _lib.GlParams()
    p.n_fft, p.win_length, p.hop, p.n_frames, p.n_iter, p.momentum = 1024, 512, 256, 60, 32, 0.99
    assert lib.neraf_griffinlim_sizes(C.byref(p), 16, C.byref(ws)) == 0
    assert ws.value >= 16 * (60 * 513 + 15104) * 4
    p.n_fft = 1000
    assert lib.neraf_griffinlim_sizes(C.byref(p), 16, C.byref(ws)) == 1
    p.n_fft, p.momentum = 1024, 1.0
    assert lib.neraf_griffinlim_sizes(C.byref(p), 16, C.byref(ws)) == 1


def test_feed_and_metrics_argument_validation_without_gpu(built):
    """The widened entry points reject bad arguments before any launch (no device needed)."""
    from neraf_b200 import _lib
    lib = _lib.lib()
    z = None
    # empty batch / no signals: a no-op, whatever the pointers
    assert lib.neraf_gather_batch(z, 4, 60, 513, z, z, z, z, 0, z, z, z, z, z, z, z, z) == 0
    assert lib.neraf_gather_batch(z, 4, 60, 513, z, z, z, z, 8, z, z, z, z, z, z, z, z) == 1
    assert b"gather_batch" in lib.neraf_last_error()
    assert lib.neraf_gather_batch(z, 4, 0, 513, z, z, z, z, 0, z, z, z, z, z, z, z, z) == 1
    mp = _lib.MetricParams()
    mp.n_samples, mp.fs, mp.t60_decay_db, mp.t60_highpass_hz = 15104, 48000.0, 10.0, 200.0
    assert lib.neraf_acoustic_metrics(C.byref(mp), z, 0, z, 0, z, z, z, z) == 0
    assert lib.neraf_acoustic_metrics(C.byref(mp), z, 4, z, 0, z, z, z, z) == 1
    mp.n_samples = 1
    assert lib.neraf_acoustic_metrics(C.byref(mp), z, 0, z, 0, z, z, z, z) == 1
    assert b"acoustic_metrics" in lib.neraf_last_error()


def test_grid_operator_argument_validation_without_gpu(built):
    """The grid-feature producer's operators reject bad windows, dtypes, strides and null pointers before any launch."""
    from neraf_b200 import _lib
    lib = _lib.lib()
    buf = (C.c_float * 64)()
    a = C.addressof(buf)
    w = _lib.Window3d(8, 8, 8, 8, 3, 1, 1)
    assert lib.neraf_grid_im2col(C.byref(w), None, 0, 8, 1, a, 0, 216, None) == 1 and b"null" in lib.neraf_last_error()
    assert lib.neraf_grid_im2col(C.byref(w), a, 0, 8, 1, a, 0, 215, None) == 1 and b"ld_col" in lib.neraf_last_error()
    assert lib.neraf_grid_im2col(C.byref(w), a, 2, 8, 1, a, 0, 216, None) == 1 and b"dtype" in lib.neraf_last_error()
    assert lib.neraf_grid_im2col(C.byref(w), a, 0, 0, 1, a, 0, 216, None) == 1 and b"strides" in lib.neraf_last_error()
    for bad in (_lib.Window3d(8, 8, 8, 8, 9, 1, 1), _lib.Window3d(8, 8, 8, 8, 3, 0, 1), _lib.Window3d(8, 8, 8, 8, 3, 1, 3),
                _lib.Window3d(0, 8, 8, 8, 3, 1, 1), _lib.Window3d(2, 2, 2, 8, 7, 1, 1), _lib.Window3d(2048, 2048, 2048, 8, 3, 1, 1)):
        assert lib.neraf_grid_col2im(C.byref(bad), a, 0, 216, a, 8, None) == 1
        assert lib.neraf_grid_maxpool(C.byref(bad), a, 0, 8, a, 8, a, None) == 1
    assert lib.neraf_grid_col2im(C.byref(w), a, 0, 216, a, 7, None) == 1 and b"row stride" in lib.neraf_last_error()
    assert lib.neraf_grid_maxpool_backward(C.byref(w), a, None, 1, 8, None, a, 8, None) == 1
    assert lib.neraf_grid_pack_weight(a, 4, 4, 27, a, 0, 107, None) == 1
    assert lib.neraf_grid_unpack_wgrad(a, 107, 4, 4, 27, 1, 0, a, None) == 1
    assert lib.neraf_grid_unpack_wgrad(a, 108, 4, 4, 27, 2, 100, a, None) == 1 and b"partial" in lib.neraf_last_error()
    assert lib.neraf_grid_unpack_wgrad(a, 108, 4, 4, 27, 0, 0, a, None) == 1
    assert lib.neraf_grid_bn_stats(a, 0, 0, 8, 8, a, None) == 1
    assert lib.neraf_grid_bn_stats(a, 0, 8, 8, 7, a, None) == 1
    assert lib.neraf_grid_bn_finalize(None, 8, 8, 1e-5, 0.1, 1, a, a, a, a, None) == 1 and b"statistics" in lib.neraf_last_error()
    assert lib.neraf_grid_bn_finalize(a, 8, 8, 1e-5, 0.1, 1, None, None, a, a, None) == 1 and b"momentum" in lib.neraf_last_error()
    assert lib.neraf_grid_bn_finalize(None, 8, 8, 1e-5, 0.0, 0, None, None, a, a, None) == 1
    assert lib.neraf_grid_bn_apply(a, 0, 8, 8, 8, a, a, a, None, None, 0, 1, a, 8, None) == 1
    assert lib.neraf_grid_bn_backward_reduce(a, None, None, a, 0, 8, 8, 8, a, a, None, a, None) == 1
    assert lib.neraf_grid_bn_backward_apply(a, a, 0, 8, 8, 8, a, a, a, None, 1, a, None, None, None) == 1
    assert lib.neraf_grid_broadcast_rows(a, 1.0, 8, 8, a, 0, 7, None) == 1


def test_product_modules_fail_loudly_without_gpu(built):
    from neraf_b200 import _lib
    from neraf_b200.field import NeRAFAudioSoundField
    from neraf_b200.griffinlim import GriffinLim
    from neraf_b200.loss import STFTLoss
    f = NeRAFAudioSoundField(1187, 512, sound_rez=1, N_frequencies=513)
    assert sorted(f.state_dict()) == sorted(
        [f"soundfield.{i}.{k}" for i in range(5) for k in ("weight", "bias")] + ["STFT_linear.0.weight", "STFT_linear.0.bias"])
    assert sum(p.numel() for p in f.parameters()) == 20428449
    with pytest.raises(_lib.NerafError):
        f(torch.zeros(2, 1187))
    with pytest.raises(_lib.NerafError):
        STFTLoss("mse")(torch.zeros(2, 1, 513), torch.zeros(2, 1, 513))
    with pytest.raises(_lib.NerafError):
        GriffinLim(n_fft=1024, win_length=512, hop_length=256, power=1)(torch.zeros(1, 513, 60))


def test_product_package_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "neraf_b200")):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


def test_tile_planner_produces_valid_schedules(built):
    """csrc/mega_plan.h on the CPU: every plan holds every tile once, can be executed in list order by spinning units
    whatever the timing (no deadlock), and beats the static stride on the field's large-batch backward."""
    out = subprocess.run([os.path.join(ROOT, "build", "mega_plan_check")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all plans valid" in out.stdout


def test_griffinlim_core_on_host_matches_torchaudio_golden(built, golden_dir):
    """Drives neraf_b200/csrc/gl_core.h (the code the GPU kernel executes) lane by lane on the CPU."""
    exe = os.path.join(ROOT, "build", "gl_host_check")
    assert subprocess.run([exe, "fft"], capture_output=True).returncode == 0
    for name in ("RAF", "SoundSpaces"):
        g = np.load(os.path.join(golden_dir, f"griffinlim_{name}.npz"))
        n, seed, n_fft, win, hop, fs = [int(v) for v in g["meta"]]
        mag = g["mag"][0, 0]
        init = (g["init_re"][0, 0] + 1j * g["init_im"][0, 0]).astype(np.complex64)
        T = mag.shape[1]
        inp = np.ascontiguousarray(mag.T, dtype=np.float32).tobytes() + np.ascontiguousarray(init.T).tobytes()
        out = subprocess.run([exe, "gl", str(n_fft), str(win), str(hop), str(T), "32", "0.99", "1"], input=inp,
                             capture_output=True)
        assert out.returncode == 0
        w = np.frombuffer(out.stdout, dtype=np.float32)
        ref = g["wave"][0, 0]
        assert np.linalg.norm(w - ref) / np.linalg.norm(ref) < 1e-4
