"""The C-ABI library loads and exports every symbol include/feba.h declares (no compute calls:
this runs without a GPU), and the host-side binding refuses to work without it."""
import ctypes
import os
import re

import pytest

import feba_b200 as fb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "feba.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(feba_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = fb.lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(fb.lib.EXPORTS) == names


def test_struct_layout_matches_header():
    # feba_settings: 16 int32 + 3 double; feba_problem: int64 + 4 int32 + 10 pointers + settings
    assert ctypes.sizeof(fb.lib.FebaSettings) == 16 * 4 + 3 * 8
    assert ctypes.sizeof(fb.lib.FebaProblem) == 8 + 4 * 4 + 10 * 8 + ctypes.sizeof(fb.lib.FebaSettings)


def test_create_fails_loudly_without_gpu_or_on_bad_settings():
    from tests import golden
    prob = golden.load_cam0()
    prob.settings.type = "fisheye"
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(fb.FebaError) as ei:
            fb.Handle(prob)
        assert ei.value.code == fb.lib.FEBA_ERR_CUDA          # no CPU fallback
    prob.settings.type = "not-a-model"
    with pytest.raises(fb.FebaError) as ei:
        fb.Handle(prob)
    assert ei.value.code == fb.lib.FEBA_ERR_INVALID and "invalid type" in ei.value.text


def test_missing_library_is_an_error(monkeypatch):
    monkeypatch.setattr(fb.lib, "_lib", None)
    monkeypatch.setattr(fb.lib, "LIB_PATH", "/nonexistent/libfeba.so")
    with pytest.raises(fb.FebaError):
        fb.lib.load()
