"""The C-ABI library builds, loads and exports exactly what include/avfe.h declares (no compute
calls here: this tier runs without a GPU)."""
import ctypes
import re
import subprocess

from conftest import ROOT


def _declared_symbols():
    text = (ROOT / "include" / "avfe.h").read_text()
    return sorted(set(re.findall(r"AVFE_API\s+[\w\s\*]+?\b(avfe_\w+)\s*\(", text)))


def test_header_symbols_are_exported(libavfe_path):
    declared = _declared_symbols()
    assert len(declared) >= 17
    out = subprocess.run(["nm", "-D", "--defined-only", str(libavfe_path)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(avfe_\w+)", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    # nothing but the documented ABI leaks out of the library
    assert exported == set(declared)


def test_ctypes_table_covers_header(libavfe_path):
    from avsl_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()
    lib = _lib.load()
    assert lib.avfe_version() >= 100
    assert lib.avfe_strerror(0) == b"ok"
    assert b"workspace" in lib.avfe_strerror(-3)
    assert lib.avfe_launch_count() == 0 or lib.avfe_launch_count() > 0


def test_library_is_sm100a_only(libavfe_path):
    out = subprocess.run(["cuobjdump", "-lelf", str(libavfe_path)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_argument_validation_without_gpu(libavfe_path):
    """Invalid-argument paths return a status before touching CUDA."""
    from avsl_b200 import _lib
    lib = _lib.load()
    assert lib.avfe_fuse(None, None, None, 0, 0.5, 0.5, 0, 4, 8, 8, None, None) == -1
    assert lib.avfe_fuse(None, None, None, 7, 0.5, 0.5, 0, 0, 8, 8, None, None) == 0      # empty batch is a no-op
    assert lib.avfe_logmel_f32(None, 1, 16000, 0, 256, None, None, None, 0, None) == -2   # n_mels > 128
    assert lib.avfe_logmel_f32(None, -1, 16000, 0, 80, None, None, None, 0, None) == -1
    assert lib.avfe_lip_roi_batch(None, 2, 1, 8, 8, None, 1, None, None, None, None, 300, 96, 88, 12,
                                  0.421, 0.165, None, None, None, None, None, None, 0, None) == -1
    assert lib.avfe_lip_workspace_bytes(10) >= 10 * 64
    assert lib.avfe_logmel_workspace_bytes(64, 480000, 0, 80) >= 64 * 4 + 80 * 8


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from avsl_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_LIB_PATH", tmp_path / "nope.so")
    import pytest
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_no_cpu_fallback_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    import avsl_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        avsl_b200.log_mel_spectrogram(np.zeros(16000, np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        avsl_b200.bgr2gray(np.zeros((4, 4, 3), np.uint8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        avsl_b200.extract_logfbank_features(np.zeros(16000, np.float32), stack_order=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        avsl_b200.spec_augment(torch.zeros(1, 80, 3000))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        avsl_b200.add_noise(np.zeros(100, np.float32), np.ones(10, np.float32), 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        avsl_b200.process_audio_for_av_hubert(np.zeros(16000, np.float32), stack_order=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        avsl_b200.fuse_transpose_layernorm(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        avsl_b200.lip_roi_collate(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), torch.zeros(2, dtype=torch.int64),
                                  torch.zeros(1, 68, 2, dtype=torch.float64), T_pad=1)


def test_argument_validation_of_the_widened_entry_points(libavfe_path):
    """Status codes of the rows added after the hot path (collation, SpecAugment masks, logfbank,
    fused LayerNorm) for invalid arguments, before anything touches CUDA."""
    from avsl_b200 import _lib
    lib = _lib.load()
    assert lib.avfe_fuse_layernorm(None, None, None, 0, 0.5, 0.5, 0, 4, 8, 8, None, None, 1e-5, None, None) == -1
    assert lib.avfe_fuse_layernorm(None, None, None, 9, 0.5, 0.5, 0, 4, 8, 8, None, None, 1e-5, None, None) == -1
    assert lib.avfe_fuse_layernorm(None, None, None, 0, 0.5, 0.5, 0, 0, 8, 8, None, None, 1e-5, None, None) == 0
    assert lib.avfe_spec_mask_f32(None, 2, 80, 3000, None, 4, 0.0, None) == -1
    assert lib.avfe_spec_mask_f32(None, 2, 80, 3000, None, 0, 0.0, None) == 0           # no bands: no-op
    assert lib.avfe_logfbank_f32(None, None, None, 1, 16000, None, 26, 3, 1, None, None, 0, None) == -2   # 8 % 3 != 0
    assert lib.avfe_logfbank_f32(None, None, None, 1, 16000, None, 26, 4, 1, None, None, 0, None) == -1
    assert lib.avfe_logfbank_num_frames(16000) == 99 and lib.avfe_logfbank_num_frames(400) == 1
    assert lib.avfe_logfbank_workspace_bytes() >= 26 * 12
    assert lib.avfe_lip_roi_collate(None, 3, 1, 8, 8, None, 1, None, None, None, None, 300, 96, 88, 12, 0.421, 0.165,
                                    None, 0, None, None, None, None, 0, None) == -1           # T_pad < 1
    assert lib.avfe_add_noise(None, None, None, None, None, 0, 100, None, None, None, 0, None) == 0     # empty batch
    assert lib.avfe_add_noise(None, None, None, None, None, 2, 100, None, None, None, 0, None) == -1
    assert lib.avfe_add_noise(None, None, None, None, None, -1, 100, None, None, None, 0, None) == -1
    # workspace: two heaps of 2^(depth+1) floats per clip; 30 s clips need depth 13
    assert lib.avfe_add_noise_workspace_bytes(64, 480000) == 256 + 64 * 2 * (2 << 13) * 4 + 64 * 16
    assert lib.avfe_add_noise_workspace_bytes(1, 128) == 256 + 2 * 2 * 4 + 16
    assert lib.avfe_add_noise_workspace_bytes(1, 1 << 40) == 0                                           # unsupported length
