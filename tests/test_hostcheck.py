"""No-GPU checks of the library's __host__ __device__ codelets (tests/hostcheck is a TEST-ONLY
CPU build of avfe_logmel_core.cuh / avfe_lip_math.cuh) against the oracle: index algebra of the
two-frames-per-FFT 20x20 STFT, reflect padding, similarity fit, float64 bilinear blend."""
import ctypes

import numpy as np
import pytest

from avsl_b200 import synth
from avsl_b200.lips import mean_face_landmarks
from oracle import lips as OL
from oracle import logmel as OM

from conftest import vp


def test_dft20(hostcheck):
    rng = np.random.default_rng(0)
    x = (rng.normal(size=20) + 1j * rng.normal(size=20)).astype(np.complex64)
    o = np.zeros(20, np.complex64)
    hostcheck.hc_dft20(vp(x), vp(o))
    assert np.abs(o - np.fft.fft(x.astype(np.complex128))).max() < 5e-6


def _raw_log10_f64(a, Lp, n_mels, frames):
    a = np.pad(a.astype(np.float64), (0, Lp - len(a)))
    n = np.arange(400)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * n / 400)
    ap = np.pad(a, (200, 200), mode="reflect")
    idx = frames[:, None] * 160 + n[None, :]
    spec = np.fft.rfft(ap[idx] * w, axis=-1)
    P = spec.real ** 2 + spec.imag ** 2
    return np.log10(np.maximum(OM.mel_filters(n_mels).astype(np.float64) @ P.T, 1e-10))


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("L,pad", [(48000, 0), (20011, 3989), (1000, 0)])
def test_logmel_tile_codelets(hostcheck, n_mels, L, pad):
    a = synth.audio_clip(L, 3) * 2.0
    Lp = L + pad
    n_frames = Lp // 160
    fb = OM.mel_filters(n_mels)
    for t0 in sorted({0, 32 * ((n_frames - 1) // 32)}):
        out = np.zeros((n_mels, 32), np.float32)
        hostcheck.hc_logmel_tile(vp(a), ctypes.c_int64(L), ctypes.c_int64(Lp), ctypes.c_int64(t0),
                                 n_mels, vp(fb), vp(out))
        nv = min(32, n_frames - t0)
        ref = _raw_log10_f64(a, Lp, n_mels, np.arange(t0, t0 + nv))
        assert np.abs(out[:, :nv] - ref).max() < 2e-5


def test_dft16_and_logfbank_codelets(hostcheck):
    """The warp-per-FFT 16 x 32 two-frames-per-FFT 512-point transform, pre-emphasis and the
    filterbank against the oracle's restatement of python_speech_features.logfbank."""
    from oracle import logfbank as OF
    rng = np.random.default_rng(1)
    x = (rng.normal(size=16) + 1j * rng.normal(size=16)).astype(np.complex64)
    o = np.zeros(16, np.complex64)
    hostcheck.hc_dft16(vp(x), vp(o))
    assert np.abs(o - np.fft.fft(x.astype(np.complex128))).max() < 2e-6
    fb = np.ascontiguousarray(OF.get_filterbanks().astype(np.float32))
    for L in (16000, 5000, 401, 400, 37):
        a = np.ascontiguousarray(synth.audio_clip(L, 7) * 3.0)
        ref = OF.logfbank(a)
        nfr = OF.num_frames(L)
        for fa in sorted({0, nfr - 1, nfr // 2}):
            fb_idx = min(fa + 1, nfr - 1)
            out = np.zeros((2, 26), np.float32)
            hostcheck.hc_logfbank_pair(vp(a), ctypes.c_int64(L), ctypes.c_int64(fa), ctypes.c_int64(fb_idx), 26,
                                       vp(fb), vp(out))
            assert np.abs(out[0] - ref[fa]).max() < 2e-4, (L, fa)
            assert np.abs(out[1] - ref[fb_idx]).max() < 2e-4, (L, fb_idx)


def test_float_key_order(hostcheck):
    hostcheck.hc_key_float.restype = ctypes.c_float
    vals = np.array([-np.inf, -10.0, -1e-3, -0.0, 0.0, 1e-10, 1.5, 3e38, np.inf], dtype=np.float32)
    keys = [hostcheck.hc_float_key(ctypes.c_float(v)) for v in vals]
    assert keys == sorted(keys)
    for v, k in zip(vals, keys):
        assert hostcheck.hc_key_float(k) == v


def test_gray_codelet(hostcheck):
    g = np.load(pytest.importorskip("pathlib").Path(__file__).parent / "golden" / "gray_golden.npz")
    sweep = np.ascontiguousarray(g["sweep"])
    out = np.zeros(sweep.shape[:2], np.uint8)
    hostcheck.hc_gray(vp(sweep), ctypes.c_int64(out.size), vp(out))
    np.testing.assert_array_equal(out, g["sweep_gray"])


def test_similarity_fit_matches_umeyama(hostcheck):
    mf = mean_face_landmarks()
    rng = np.random.default_rng(5)
    for _ in range(50):
        src = mf[OL.STABLE_IDS] * rng.uniform(0.5, 1.5) + rng.normal(0, 3, size=(5, 2)) + rng.uniform(-20, 20, 2)
        dst = np.ascontiguousarray(mf[OL.STABLE_IDS])
        src = np.ascontiguousarray(src)
        fwd, inv = np.zeros(6), np.zeros(6)
        hostcheck.hc_similarity_fit(vp(src), vp(dst), 5, vp(fwd), vp(inv))
        T = OL.umeyama(src, dst, True)
        np.testing.assert_allclose(fwd.reshape(2, 3), T[:2], rtol=1e-12, atol=1e-11)
        np.testing.assert_allclose(inv.reshape(2, 3), np.linalg.inv(T)[:2], rtol=1e-12, atol=1e-11)
    # reflection-like configuration (det(A) < 0): still the proper similarity skimage returns
    src = np.ascontiguousarray(mf[OL.STABLE_IDS] * np.array([-1.0, 1.0]) + 300)
    fwd, inv = np.zeros(6), np.zeros(6)
    hostcheck.hc_similarity_fit(vp(src), vp(np.ascontiguousarray(mf[OL.STABLE_IDS])), 5, vp(fwd), vp(inv))
    np.testing.assert_allclose(fwd.reshape(2, 3), OL.umeyama(src, mf[OL.STABLE_IDS])[:2], rtol=1e-10, atol=1e-9)


def test_crop_origin_codelet(hostcheck):
    rc = np.zeros(2, np.int32)
    for cx, cy in [(10.0, 290.0), (100.5, 150.5), (101.5, 151.5), (129.31, 157.82), (0.0, 0.0), (299.9, 10.2)]:
        hostcheck.hc_crop_origin(ctypes.c_double(cx), ctypes.c_double(cy), 48, 300, vp(rc))
        assert tuple(rc) == OL.cut_patch_origin(np.array([[cx, cy]] * 20), 48, 48, (300, 300))
    hostcheck.hc_crop_origin(ctypes.c_double(np.nan), ctypes.c_double(1.0), 48, 300, vp(rc))
    assert tuple(rc) == (-1, -1)


def test_warp_window_bit_exact_vs_oracle(hostcheck):
    """Same inverse matrix in, identical uint8 out — including windows hanging over the frame
    border (cval 0 taps) and flat saturated regions (the k/255 round-trip cases)."""
    frames, lm, _ = synth.video_clip(3, 224, 224, seed=11, invalid_frac=0.0)
    gray = OL.bgr2gray(frames)
    gray[1, 60:140, 60:140] = 255        # flat regions: blend of equal taps
    gray[2, 60:140, 60:140] = 77
    mf = mean_face_landmarks()
    for f in range(3):
        T = OL.umeyama(lm[f][OL.STABLE_IDS], mf[OL.STABLE_IDS])
        M = np.linalg.inv(T)
        for (r0, c0) in [(110, 81), (0, 0), (204, 204), (150, 20)]:
            ref = OL.to_u8(OL.warp_float(gray[f], M, (300, 300), (r0, r0 + 96), (c0, c0 + 96)))
            out = np.zeros((96, 96), np.uint8)
            inv6 = np.ascontiguousarray(M[:2].reshape(-1))
            hostcheck.hc_warp_window(vp(np.ascontiguousarray(gray[f])), 224, 224, vp(inv6), r0, c0, 96, 96, vp(out))
            np.testing.assert_array_equal(out, ref)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 128, 129, 136, 143, 144, 255, 257, 272, 1000, 4097, 16000, 70001, 480000])
def test_noise_tree_sum_is_numpys_pairwise_sum(hostcheck, n):
    """The heap-addressed tree of avfe_noise.cu adds in numpy's order: same float32 bits as np.sum of
    the squares, for the plain signal and for a tiled (i mod period) one, at a launch depth sized
    for a longer clip as well."""
    hostcheck.hc_noise_tree_sumsq.restype = ctypes.c_float
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 3000).astype(np.float32)
    leaves = ctypes.c_int(0)
    for max_len in (n, 2 * n + 77):
        got = hostcheck.hc_noise_tree_sumsq(vp(x), ctypes.c_uint32(n), ctypes.c_uint32(n), ctypes.c_int64(max_len),
                                            ctypes.byref(leaves))
        assert np.float32(got) == np.sum(np.square(x)), (n, max_len)
    period = max(1, n // 3 + 1)
    tiled = x[np.arange(n) % period]
    got = hostcheck.hc_noise_tree_sumsq(vp(x), ctypes.c_uint32(n), ctypes.c_uint32(period), ctypes.c_int64(n),
                                        ctypes.byref(leaves))
    assert np.float32(got) == np.sum(np.square(tiled))
    assert leaves.value >= max(1, n // 128)
