"""Oracle for the AV-HuBERT audio features (CPU tier): self-consistency of the restatement of
python_speech_features.logfbank (the package is absent here: parity unpinned, see oracle/logfbank.py)."""
import numpy as np

from avsl_b200 import synth
from avsl_b200.audio import _logfbank_filters_np
from oracle import logfbank as OF


def test_filterbank_shape_and_partition():
    fb = OF.get_filterbanks()
    assert fb.shape == (26, 257) and fb.min() >= 0.0 and fb.max() <= 1.0
    # neighbouring triangles overlap so that interior bins sum to one
    s = fb.sum(axis=0)
    inner = slice(int(np.nonzero(fb[0])[0].max()) + 1, int(np.nonzero(fb[-1])[0].min()))
    np.testing.assert_allclose(s[inner], 1.0, atol=1e-12)
    np.testing.assert_array_equal(_logfbank_filters_np(), fb.astype(np.float32))


def test_frame_count_and_padding_rule():
    assert [OF.num_frames(n) for n in (1, 400, 401, 560, 561, 16000)] == [1, 1, 2, 2, 3, 99]
    a = synth.audio_clip(1000, 3)
    f = OF.logfbank(a)
    assert f.shape == (OF.num_frames(1000), 26) and f.dtype == np.float64
    assert np.allclose(OF.logfbank(np.zeros(800, np.float32)), np.log(np.finfo(float).eps))


def test_stack_and_normalise():
    a = synth.audio_clip(16000 * 2 + 300, 4)
    f1 = OF.extract_logfbank_features(a, stack_order=1)
    f4 = OF.extract_logfbank_features(a, stack_order=4)
    assert f1.dtype == np.float32 and f4.shape == (-(-len(f1) // 4), 104)
    np.testing.assert_array_equal(f4[0], f1[:4].reshape(-1))
    pad = 4 * len(f4) - len(f1)
    assert pad and not f4[-1, -26 * pad:].any()
    n = OF.audio_to_tensor(f4)
    assert np.abs(n.mean(axis=1)).max() < 1e-5 and np.abs(n.std(axis=1) - 1.0).max() < 1e-3
