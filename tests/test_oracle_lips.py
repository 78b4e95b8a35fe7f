"""Checks of the lip-path oracle (oracle/lips.py): pinned parts against cv2 and the reference's
own function (golden), unpinned parts (similarity fit, warp) against independent libraries."""
import numpy as np
import pytest

from avsl_b200 import synth
from avsl_b200.lips import mean_face_landmarks
from oracle import lips as O

from conftest import GOLDEN


def test_gray_matches_cv2_golden():
    g = np.load(GOLDEN / "gray_golden.npz")
    np.testing.assert_array_equal(O.bgr2gray(g["img"]), g["img_gray"])
    np.testing.assert_array_equal(O.bgr2gray(g["sweep"]), g["sweep_gray"])


def test_gray_matches_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(120, 77, 3), dtype=np.uint8)
    np.testing.assert_array_equal(O.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


def test_video_feats_matches_reference_function_golden():
    g = np.load(GOLDEN / "video_feats_golden.npz")
    out = O.video_feats_from_u8(g["roi"])
    assert out.dtype == np.float32 and out.shape == (3, 88, 88, 1)
    np.testing.assert_array_equal(out, g["feats"])
    np.testing.assert_array_equal(O.video_feats_from_u8(g["levels"]), g["levels_feats"])


def test_landmarks_interpolate_semantics():
    a, b = np.full((68, 2), 10, dtype=np.int32), np.full((68, 2), 20, dtype=np.int32)
    out = O.landmarks_interpolate([None, a, None, None, None, b, None])
    assert out[0] is a and out[6] is b          # ends replicate
    np.testing.assert_allclose(out[2], 10 + 1 / 4 * 10)
    np.testing.assert_allclose(out[4], 10 + 3 / 4 * 10)
    assert O.landmarks_interpolate([None, None]) is None
    assert O.landmarks_interpolate([a])[0] is a


def test_umeyama_recovers_known_similarity():
    rng = np.random.default_rng(1)
    src = rng.normal(size=(5, 2)) * 30 + 100
    th, s, t = 0.3, 1.7, np.array([12.0, -7.0])
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    dst = s * src @ R.T + t
    T = O.umeyama(src, dst, True)
    np.testing.assert_allclose(T[:2, :2], s * R, atol=1e-12)
    np.testing.assert_allclose(T[:2, 2], t, atol=1e-10)
    np.testing.assert_array_equal(T[2], [0, 0, 1])


def test_warp_matches_scipy_interior_and_cv2():
    ndi = pytest.importorskip("scipy.ndimage")
    cv2 = pytest.importorskip("cv2")
    frames, lm, _ = synth.video_clip(2, 224, 224, seed=4)
    gray = O.bgr2gray(frames[0])
    mf = mean_face_landmarks()
    tf = O.SimilarityTransform(O.umeyama(lm[0][O.STABLE_IDS], mf[O.STABLE_IDS]))
    M = tf.inverse.params
    w = O.warp_float(gray, M, (300, 300))
    # independent bilinear: scipy map_coordinates, grid-constant == skimage 'constant'
    rr, cc = np.meshgrid(np.arange(300.0), np.arange(300.0), indexing="ij")
    x = M[0, 0] * cc + M[0, 1] * rr + M[0, 2]
    y = M[1, 0] * cc + M[1, 1] * rr + M[1, 2]
    ref = ndi.map_coordinates(gray.astype(np.float64) / 255.0, [y, x], order=1, mode="grid-constant", cval=0.0)
    assert np.abs(w - ref).max() < 1e-12
    # cv2.warpAffine (fixed-point coordinates): agreement to ~1 grey level in the interior
    cvw = cv2.warpAffine(gray, tf.params[:2], (300, 300), flags=cv2.INTER_LINEAR)
    diff = np.abs(O.to_u8(w).astype(int) - cvw.astype(int))[20:-20, 20:-20]
    assert np.percentile(diff, 99) <= 2


def test_window_restriction_equals_full_warp():
    frames, lm, valid = synth.video_clip(14, 160, 160, seed=7)
    gray = O.bgr2gray(frames)
    lms = [lm[i] if valid[i] else None for i in range(len(lm))]
    mf = mean_face_landmarks()
    a, ta, oa = O.extract_lip_frames_from_arrays(gray, lms, mf, full_warp=False)
    b, tb, ob = O.extract_lip_frames_from_arrays(gray[:3], lms[:3], mf, full_warp=True)
    assert a.shape == (14, 96, 96) and a.dtype == np.uint8
    np.testing.assert_array_equal(O.extract_lip_frames_from_arrays(gray[:3], lms[:3], mf)[0], b)


def test_window_semantics_short_and_tail():
    """T < 12: one transform from the mean of all frames, reused for every frame;
    T >= 12: frames after T-12 reuse the transform of frame T-12."""
    mf = mean_face_landmarks()
    for T in (1, 7, 12, 15):
        frames, lm, _ = synth.video_clip(T, 128, 128, seed=T, invalid_frac=0.0)
        gray = O.bgr2gray(frames)
        _, tf, _ = O.extract_lip_frames_from_arrays(gray, list(lm), mf)
        margin = min(T, 12)
        for i in range(T):
            j = min(i, T - margin)
            exp = O.umeyama(lm[j:j + margin].mean(axis=0)[O.STABLE_IDS], mf[O.STABLE_IDS])
            np.testing.assert_allclose(tf[i], exp, rtol=1e-13, atol=1e-13)


def test_cut_patch_clamps_and_rounds():
    img = np.arange(300 * 300, dtype=np.int64).reshape(300, 300).astype(np.uint8)
    lm = np.array([[10.0, 290.0]] * 20)           # (x, y): far left, far bottom -> clamped
    assert O.cut_patch_origin(lm, 48, 48, img.shape) == (204, 0)
    lm = np.array([[100.5, 150.5]] * 20)          # round half to even: 100, 150
    assert O.cut_patch_origin(lm, 48, 48, img.shape) == (102, 52)
    lm = np.array([[101.5, 151.5]] * 20)          # -> 102, 152
    assert O.cut_patch_origin(lm, 48, 48, img.shape) == (104, 54)
    assert O.cut_patch(img, lm, 48, 48).shape == (96, 96)


def test_no_detection_returns_empty():
    out, tf, org = O.extract_lip_frames_from_arrays(np.zeros((3, 64, 64), np.uint8), [None] * 3,
                                                     mean_face_landmarks())
    assert out.size == 0 and tf is None


@pytest.mark.parametrize("kind", ["translate", "border", "scale2"])
def test_oracle_warp_known_answers(kind):
    """The restated skimage warp + cut_patch against reference-independent known answers (integer
    translation -> source pixels, border-straddling ROI -> exact zeros, x2 zoom -> truncated
    midpoints): the only check of the unpinned warp chain that does not compare two restatements."""
    from known_answers import known_answer_case
    frames, gray, lm, tf, expect, (r0, c0) = known_answer_case(kind)
    fwd, inv = tf[0, :9].reshape(3, 3), tf[0, 9:].reshape(3, 3)
    assert O.cut_patch_origin(O.SimilarityTransform(fwd)(lm[0])[48:68], 48, 48, (300, 300)) == (r0, c0)
    for t in range(len(frames)):
        full = O.apply_transform(O.SimilarityTransform(fwd), gray[t])
        np.testing.assert_array_equal(full[r0:r0 + 96, c0:c0 + 96], expect[t])


def test_video_feats_rgb_dark_resize_match_reference_function_golden():
    """utils/hf_video_utils.py:103-138 on decord-style input: goldens are the reference function's
    own outputs (make_golden.py::video_feats)."""
    g = np.load(GOLDEN / "video_feats_golden.npz")
    np.testing.assert_array_equal(O.video_feats_from_frames(g["rgb"]), g["rgb_feats"])
    np.testing.assert_array_equal(O.video_feats_from_frames(g["dark"]), g["dark_feats"])
    np.testing.assert_array_equal(O.video_feats_from_frames(g["roi"][..., None]), g["feats"])
    for name in ("small_rgb", "small_gray", "small_tall", "small_wide"):
        got = O.video_feats_from_frames(g[name])
        np.testing.assert_array_equal(got, g[name + "_feats_noipp"])          # cv2's own code path
        assert np.abs(got - g[name + "_feats_ipp"]).max() <= 5e-5              # cv2 as shipped on x86 (IPP)


def test_resize_restatement_against_live_cv2():
    import cv2
    rng = np.random.default_rng(5)
    was = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        for shape in [(64, 64), (87, 88), (40, 200), (120, 33)]:
            for dt in (np.float32, np.float64):
                src = (rng.integers(0, 256, shape) / 255.0).astype(dt)
                np.testing.assert_array_equal(O.resize_linear(src, 88), cv2.resize(src, (88, 88)))
    finally:
        cv2.ipp.setUseIPP(was)


def test_rgb_dot_float32_cast_is_order_independent():
    """All 2^24 RGB triples: float32(np.dot(rgb, w)) (BLAS order, FMA or not) equals float32 of the
    plain left-to-right float64 sum the oracle and the kernel use."""
    v = np.arange(256, dtype=np.uint8)
    r, gg, b = np.meshgrid(v, v, v, indexing="ij")
    rgb = np.stack([r, gg, b], axis=-1).reshape(-1, 3)
    ref = np.dot(rgb, [0.2989, 0.5870, 0.1140]).astype(np.float32)
    f = rgb.astype(np.float64)
    mine = ((f[:, 0] * 0.2989 + f[:, 1] * 0.5870) + f[:, 2] * 0.1140).astype(np.float32)
    np.testing.assert_array_equal(mine, ref)
