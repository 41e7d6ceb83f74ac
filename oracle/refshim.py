"""Make the reference's own pure-torch classes importable in the BUILD container (test infrastructure only).

/root/reference/NeRAF/NeRAF_field.py:10-25 and NeRAF_helper.py:4 import
nerfstudio / pyroomacoustics at module top; neither is installed.  Registering
empty stand-in modules lets the REAL ``NeRAFAudioSoundField``, ``STFTLoss``,
``measure_edt`` and ``measure_clarity`` run unmodified.  /root/reference does not
exist on the GPU box, so this is used only by oracle/make_golden.py and by tests
that skip when the reference tree is absent.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "NeRAF"))


def install() -> None:
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")

    class _Dummy:  # placeholder for imported-but-unused names
        def __init__(self, *a, **k):
            pass

    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("nerfstudio")
    mod("nerfstudio.field_components")
    mod("nerfstudio.field_components.spatial_distortions", SpatialDistortion=_Dummy)
    mod("nerfstudio.field_components.encodings", NeRFEncoding=_Dummy, SHEncoding=_Dummy)
    mod("nerfstudio.fields")
    mod("nerfstudio.fields.nerfacto_field", NerfactoField=_Dummy)
    mod("nerfstudio.fields.base_field", Field=_Dummy)
    mod("nerfstudio.model_components")
    mod("nerfstudio.model_components.losses", MSELoss=_Dummy, distortion_loss=None, interlevel_loss=None,
        orientation_loss=None, pred_normal_loss=None, scale_gradients_by_distance_squared=None)
    pra = mod("pyroomacoustics")
    pra.experimental = types.SimpleNamespace(measure_rt60=None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """Returns (NeRAFAudioSoundField, STFTLoss, helper module) -- the reference's real objects."""
    install()
    from NeRAF.NeRAF_field import NeRAFAudioSoundField
    from NeRAF.NeRAF_evaluator import STFTLoss
    import NeRAF.NeRAF_helper as helper
    return NeRAFAudioSoundField, STFTLoss, helper


def load_datasets():
    """Returns the reference's real (RAFDataset, SoundSpacesDataset) classes (NeRAF_dataset.py).

    NeRAF_dataparser.py:16,21 needs nerfstudio's dataparser base classes (only subclassed, never called by the
    datasets) and NeRAF_dataset.py:21 imports librosa (used by the RAF wav loader only): stand-ins for both, so the
    SoundSpaces ``get_data`` path -- np.load of a magnitude file, one column, log, poses -- runs unmodified.
    """
    install()
    import dataclasses

    class _Cfg:
        pass

    @dataclasses.dataclass
    class _Outputs:
        pass

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    for name in ("nerfstudio", "nerfstudio.data", "nerfstudio.data.dataparsers"):
        mod(name).__path__ = []          # make them packages
    mod("nerfstudio.data.dataparsers.base_dataparser", DataParser=_Cfg, DataParserConfig=_Cfg, DataparserOutputs=_Outputs)
    mod("nerfstudio.data.scene_box", SceneBox=_Cfg)
    if "librosa" not in sys.modules:
        try:
            import librosa  # noqa: F401
        except ImportError:
            mod("librosa")
    from NeRAF.NeRAF_dataset import RAFDataset, SoundSpacesDataset
    return RAFDataset, SoundSpacesDataset
