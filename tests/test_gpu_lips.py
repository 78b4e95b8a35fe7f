"""GPU parity: the lip-ROI path through the C ABI vs the oracle.

Bars (north_star / BASELINE.md 5): gray frames and crop indices bit-exact; ROI pixels within
1 grey level (normalised: (1/255)/0.165) with the device-fitted transform and BIT-EXACT when the
transform matrices are supplied (the float64 blend follows skimage's operation order)."""
import numpy as np
import pytest
import torch

import avsl_b200 as A
from avsl_b200 import lips as L
from avsl_b200 import synth
from oracle import lips as O

from conftest import GOLDEN
from known_answers import known_answer_case

pytestmark = pytest.mark.gpu


def _as_list(lm, valid):
    return [lm[i] if valid[i] else None for i in range(len(lm))]


def test_gray_bit_exact_golden_and_shapes():
    g = np.load(GOLDEN / "gray_golden.npz")
    np.testing.assert_array_equal(A.bgr2gray(g["img"]), g["img_gray"])
    np.testing.assert_array_equal(A.bgr2gray(g["sweep"]), g["sweep_gray"])
    rng = np.random.default_rng(0)
    for shape in [(5, 224, 224, 3), (3, 37, 53, 3), (1, 1, 1, 3), (2, 288, 352, 3), (0, 8, 8, 3)]:
        x = rng.integers(0, 256, size=shape, dtype=np.uint8)
        np.testing.assert_array_equal(A.bgr2gray(x), O.bgr2gray(x))
    # misaligned device pointer -> scalar kernel
    x = torch.from_numpy(rng.integers(0, 256, size=(1 + 4 * 64 * 64 * 3,), dtype=np.uint8)).cuda()
    v = x[1:].view(4, 64, 64, 3)
    np.testing.assert_array_equal(A.bgr2gray(v).cpu().numpy(), O.bgr2gray(v.cpu().numpy()))


def test_video_feats_golden_reference_function():
    g = np.load(GOLDEN / "video_feats_golden.npz")
    out = A.load_video_feats(g["roi"])
    assert out.dtype == np.float32 and out.shape == (3, 88, 88, 1)
    np.testing.assert_array_equal(out, g["feats"])
    np.testing.assert_array_equal(A.load_video_feats(g["levels"]), g["levels_feats"])
    with pytest.raises(ValueError, match="Expected 3D frames"):
        A.load_video_feats(np.zeros((2, 2, 96, 96, 3), np.uint8))


def test_video_feats_as_the_call_site_runs_it():
    """utils/hf_video_utils.py:103-138 on what decord returns: RGB frames (float64 dot, /255 if the
    stack is not all-dark), and frames smaller than the crop (cv2.resize).  Goldens are the
    reference function's own outputs (tests/golden/make_golden.py::video_feats)."""
    g = np.load(GOLDEN / "video_feats_golden.npz")
    np.testing.assert_array_equal(A.load_video_feats(g["rgb"]), g["rgb_feats"])
    np.testing.assert_array_equal(A.load_video_feats(g["dark"]), g["dark_feats"])
    # CUDA tensor in -> CUDA tensor out, same numbers
    t = A.load_video_feats(torch.from_numpy(g["rgb"]).cuda())
    assert t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == (2, 88, 88, 1)
    np.testing.assert_array_equal(t.cpu().numpy(), g["rgb_feats"])
    for name in ("small_rgb", "small_gray", "small_tall", "small_wide"):
        got = A.load_video_feats(g[name])
        assert got.shape == g[name + "_feats_ipp"].shape and got.dtype == np.float32
        # bit-exact against cv2's own code (IPP off) through the reference function ...
        np.testing.assert_array_equal(got, g[name + "_feats_noipp"])
        np.testing.assert_array_equal(got, O.video_feats_from_frames(g[name]))
        # ... and within 5e-5 (normalised units) of cv2 as shipped on x86 (IPP's float32 resize)
        assert np.abs(got - g[name + "_feats_ipp"]).max() <= 5e-5


def test_video_feats_rgb_every_triple():
    """All 2^24 RGB triples in one 4096 x 4096 frame: float32(dot)/255, normalised -- bit-exact
    against numpy (the crop equals the frame, so nothing is cut)."""
    v = np.arange(256, dtype=np.uint8)
    r, gg, b = np.meshgrid(v, v, v, indexing="ij")
    frame = np.stack([r, gg, b], axis=-1).reshape(1, 4096, 4096, 3)
    got = A.load_video_feats(frame, image_crop_size=4096)
    ref = ((np.dot(frame[..., :3], [0.2989, 0.5870, 0.1140]).astype(np.float32) / 255.0 - 0.421) / 0.165)
    np.testing.assert_array_equal(got[..., 0], ref.astype(np.float32))


@pytest.mark.parametrize("kind", ["translate", "border", "scale2"])
@pytest.mark.parametrize("as_gray", [False, True])
def test_known_answer_warp(kind, as_gray):
    """Reference-independent known answers through avfe_lip_roi_batch (tforms_in): an integer
    translation returns source pixels exactly, a border-straddling ROI exact zeros outside the
    frame, a x2 zoom the truncated midpoints -- the expected bytes come from index arithmetic and
    the four-pixel average only (no coordinate transform, no interpolation routine).  BGR input
    takes the frame-owner kernel, gray input the generic one."""
    frames, gray, lm, tf, expect, (r0, c0) = known_answer_case(kind)
    T = len(frames)
    src = torch.from_numpy(gray if as_gray else frames).cuda()
    res = L.lip_roi_batch(src, torch.tensor([0, T], dtype=torch.int64).cuda(), torch.from_numpy(lm).cuda(),
                          torch.ones(T, dtype=torch.uint8).cuda(), tforms_in=torch.from_numpy(tf).cuda(),
                          want_gray=not as_gray, want_u8=True, want_f32=True, want_meta=True)
    np.testing.assert_array_equal(res.crop_rc.cpu().numpy(), np.tile(np.array([[r0, c0]], np.int32), (T, 1)))
    np.testing.assert_array_equal(res.lip_u8.cpu().numpy(), expect)
    if kind == "border":
        assert (expect[:, :, :20] == 0).all() and (expect[:, :10, :] == 0).all() and expect[:, 60:, 60:].any()
    np.testing.assert_array_equal(res.lip_f32.cpu().numpy(), O.video_feats_from_u8(expect)[..., 0])
    if not as_gray:
        np.testing.assert_array_equal(res.gray.cpu().numpy(), gray)


def test_landmarks_interpolate_bit_exact():
    _, lm, valid = synth.video_clip(1, 8, 8)       # frames unused
    lm, valid = synth.landmarks_for_clip(60, seed=9, invalid_frac=0.4)
    valid[:3] = 0
    valid[-2:] = 0
    lst = _as_list(lm.astype(np.int32), valid)
    got = A.landmarks_interpolate(list(lst))
    ref = O.landmarks_interpolate(list(lst))
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))
    assert A.landmarks_interpolate([None, None]) is None


def test_warp_img_apply_transform_cut_patch():
    frames, lm, _ = synth.video_clip(2, 224, 224, seed=4, invalid_frac=0.0)
    gray = O.bgr2gray(frames)
    mf = A.mean_face_landmarks()
    warped, tform = A.warp_img(lm[0][O.STABLE_IDS], mf[O.STABLE_IDS], gray[0], (300, 300))
    ref_w, ref_t = O.warp_img(lm[0][O.STABLE_IDS], mf[O.STABLE_IDS], gray[0], (300, 300))
    np.testing.assert_allclose(tform.params, ref_t.params, rtol=1e-12, atol=1e-10)
    d = np.abs(warped.astype(int) - ref_w.astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 1e-3
    # a GIVEN transform is warped bit-exactly (apply_transform semantics)
    np.testing.assert_array_equal(A.apply_transform(ref_t, gray[1], (300, 300)), O.apply_transform(ref_t, gray[1]))
    t_lm = ref_t(lm[0])
    np.testing.assert_array_equal(A.cut_patch(ref_w, t_lm[48:68], 48, 48), O.cut_patch(ref_w, t_lm[48:68], 48, 48))
    edge = np.array([[5.0, 298.0]] * 20)
    np.testing.assert_array_equal(A.cut_patch(ref_w, edge, 48, 48), O.cut_patch(ref_w, edge, 48, 48))
    np.testing.assert_allclose(tform(lm[0]), t_lm, rtol=1e-10)


@pytest.mark.parametrize("T,H,W,inv", [(30, 224, 224, 0.05), (7, 224, 224, 0.2), (1, 224, 224, 0.0),
                                       (12, 160, 200, 0.0), (13, 288, 352, 0.1)])
def test_extract_lip_frames_vs_oracle(T, H, W, inv):
    frames, lm, valid = synth.video_clip(T, H, W, seed=100 + T, invalid_frac=inv)
    lst = _as_list(lm, valid)
    got = A.extract_lip_frames(frames, lst)
    ref, tf, org = O.extract_lip_frames_from_arrays(O.bgr2gray(frames), lst, A.mean_face_landmarks())
    assert got.shape == ref.shape == (T, 96, 96) and got.dtype == np.uint8
    d = np.abs(got.astype(int) - ref.astype(int))
    assert d.max() <= 1
    assert (d != 0).mean() < 1e-3
    # gray input (already converted) gives the same ROI
    np.testing.assert_array_equal(A.extract_lip_frames(O.bgr2gray(frames), lst), got)


def test_fused_batch_config2_outputs():
    """Config 2 shape (10 s, 25 fps, 224x224) plus a T=7 clip in the same batch: gray bit-exact,
    crop indices bit-exact, transforms to 1e-12, f32 features within one grey level."""
    clips = [synth.video_clip(250, 224, 224, seed=3407), synth.video_clip(7, 224, 224, seed=8, invalid_frac=0.3)]
    frames = np.concatenate([c[0] for c in clips])
    lm = np.concatenate([c[1] for c in clips])
    valid = np.concatenate([c[2] for c in clips])
    off = torch.tensor([0, 250, 257], dtype=torch.int64).cuda()
    res = L.lip_roi_batch(torch.from_numpy(frames).cuda(), off, torch.from_numpy(lm).cuda(),
                          torch.from_numpy(valid).cuda(), want_gray=True, want_u8=True, want_f32=True,
                          want_meta=True)
    gray_ref = O.bgr2gray(frames)
    np.testing.assert_array_equal(res.gray.cpu().numpy(), gray_ref)
    mf = A.mean_face_landmarks()
    lo = 0
    worst = 0.0
    for c in clips:
        T = len(c[0])
        ref, tf, org = O.extract_lip_frames_from_arrays(gray_ref[lo:lo + T], _as_list(c[1], c[2]), mf)
        np.testing.assert_array_equal(res.crop_rc[lo:lo + T].cpu().numpy(), org)
        got_tf = res.tforms[lo:lo + T].cpu().numpy()
        np.testing.assert_allclose(got_tf[:, :9].reshape(T, 3, 3), tf, rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(got_tf[:, 9:].reshape(T, 3, 3), np.linalg.inv(tf), rtol=1e-12, atol=1e-10)
        d = np.abs(res.lip_u8[lo:lo + T].cpu().numpy().astype(int) - ref.astype(int))
        assert d.max() <= 1 and (d != 0).mean() < 1e-3
        feats_ref = O.video_feats_from_u8(ref)[..., 0]
        worst = max(worst, np.abs(res.lip_f32[lo:lo + T].cpu().numpy() - feats_ref).max())
        lo += T
    assert worst <= (1.0 / 255.0) / 0.165 + 1e-6
    # f32 features are exactly the normalised u8 ROI the same launch produced
    np.testing.assert_array_equal(res.lip_f32.cpu().numpy(),
                                  O.video_feats_from_u8(res.lip_u8.cpu().numpy())[..., 0])


def test_given_transforms_are_bit_exact():
    """apply_transform semantics in the fused op: with the oracle's matrices supplied, ROI u8 and
    normalised f32 are bit-identical to the oracle."""
    frames, lm, valid = synth.video_clip(40, 224, 224, seed=21, invalid_frac=0.1)
    frames[5, 50:170, 50:170] = 255                     # flat saturated block inside the ROI footprint
    gray_ref = O.bgr2gray(frames)
    lst = _as_list(lm, valid)
    ref, tf, org = O.extract_lip_frames_from_arrays(gray_ref, lst, A.mean_face_landmarks())
    t18 = np.concatenate([tf.reshape(-1, 9), np.linalg.inv(tf).reshape(-1, 9)], axis=1)
    off = torch.tensor([0, 40], dtype=torch.int64).cuda()
    for src in (torch.from_numpy(frames).cuda(), torch.from_numpy(gray_ref).cuda()):
        res = L.lip_roi_batch(src, off, torch.from_numpy(lm).cuda(), torch.from_numpy(valid).cuda(),
                              tforms_in=torch.from_numpy(t18).cuda(), want_gray=False, want_u8=True,
                              want_f32=True, want_meta=True)
        np.testing.assert_array_equal(res.crop_rc.cpu().numpy(), org)
        np.testing.assert_array_equal(res.lip_u8.cpu().numpy(), ref)
        np.testing.assert_array_equal(res.lip_f32.cpu().numpy(), O.video_feats_from_u8(ref)[..., 0])


def test_failure_conventions():
    frames, lm, valid = synth.video_clip(5, 64, 64, seed=1)
    assert A.extract_lip_frames(frames, [None] * 5).size == 0          # no detection -> empty array
    assert A.extract_lip_frames(np.zeros((0, 64, 64, 3), np.uint8), []).size == 0
    assert A.extract_lip_frames(frames, [None] * 4).size == 0          # length mismatch is swallowed
    # ROI hanging over the frame border reads cval 0 (tiny frame, face far outside)
    far = [lm[i] + 500.0 for i in range(5)]
    got = A.extract_lip_frames(frames, far)
    ref, _, _ = O.extract_lip_frames_from_arrays(O.bgr2gray(frames), far, A.mean_face_landmarks())
    assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1


def test_idempotence_and_clip_independence_full_size():
    """Size-independent properties at BASELINE size: the same clip gives the same bytes wherever it
    sits in a batch (clips are independent), and a second pass reproduces the first."""
    f, lm, valid = synth.video_clip(250, 224, 224, seed=3407)
    F = torch.from_numpy(np.concatenate([f, f[:60], f])).cuda()
    LM = torch.from_numpy(np.concatenate([lm, lm[:60], lm])).cuda()
    V = torch.from_numpy(np.concatenate([valid, valid[:60], valid])).cuda()
    off = torch.tensor([0, 250, 310, 560], dtype=torch.int64).cuda()
    r1 = L.lip_roi_batch(F, off, LM, V, want_u8=True)
    a, b = r1.lip_f32[:250].clone(), r1.lip_f32[310:].clone()
    assert torch.equal(a, b) and torch.equal(r1.gray[:250], r1.gray[310:])
    r2 = L.lip_roi_batch(F, off, LM, V, want_u8=True)
    assert torch.equal(r2.lip_f32[:250], a) and torch.equal(r2.lip_u8, r1.lip_u8)


@pytest.mark.parametrize("shape,gray_in", [((224, 224), False), ((224, 224), True), ((120, 168), False)])
def test_collate_matches_trim_plus_collator(shape, gray_in):
    """avfe_lip_roi_collate == lip_roi_batch followed by the reference's trim
    (whisper_flamingo_ft_ami.py:299-302) and the collator's zero padding + mask, on the frame-owner
    kernel (224x224 BGR), on gray input and on a row length the bulk copies cannot take."""
    from oracle import collate as OC
    H, W = shape
    lens = [30, 9, 1, 17]
    clips = [synth.video_clip(t, H, W, seed=50 + t, invalid_frac=0.1) for t in lens]
    frames = np.concatenate([c[0] for c in clips])
    if gray_in:
        frames = O.bgr2gray(frames)
    lm = np.concatenate([c[1] for c in clips])
    valid = np.concatenate([c[2] for c in clips])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    keep = np.array([12, 9, 5, 0], dtype=np.int64)       # trimmed, untouched, longer than the clip, dropped
    d = dict(frames=torch.from_numpy(frames).cuda(), off=torch.from_numpy(off).cuda(),
             lm=torch.from_numpy(lm).cuda(), valid=torch.from_numpy(valid).cuda())
    packed = L.lip_roi_batch(d["frames"], d["off"], d["lm"], d["valid"], want_gray=not gray_in)
    feats = packed.lip_f32.cpu().numpy()
    kept = [feats[off[i]:off[i] + min(lens[i], keep[i])][..., None] for i in range(len(lens))]
    for T_pad in (12, 16):
        ref = OC.collate_video(kept, T_pad=T_pad)
        got = L.lip_roi_collate(d["frames"], d["off"], d["lm"], d["valid"], T_pad=T_pad,
                                keep_frames=torch.from_numpy(keep).cuda(), want_gray=not gray_in)
        assert got["video"].shape == (4, 1, T_pad, 88, 88) and got["padding_mask"].dtype == torch.bool
        np.testing.assert_array_equal(got["video"].cpu().numpy(), ref["video"])
        np.testing.assert_array_equal(got["padding_mask"].cpu().numpy(), ref["padding_mask"])
        if not gray_in:
            assert torch.equal(got["gray"], packed.gray)
    # no trim: every frame kept, T_pad = longest clip (the collator's own choice)
    ref = OC.collate_video([feats[off[i]:off[i + 1]][..., None] for i in range(len(lens))])
    got = L.lip_roi_collate(d["frames"], d["off"], d["lm"], d["valid"], T_pad=max(lens), want_gray=False)
    np.testing.assert_array_equal(got["video"].cpu().numpy(), ref["video"])
    np.testing.assert_array_equal(got["padding_mask"].cpu().numpy(), ref["padding_mask"])


def test_frame_kernel_edge_cases_in_one_batch():
    """The frame-owner kernel (BGR, gray wanted) on a batch that mixes: a normal clip, a clip with
    no detection at all (zero ROIs, crop_rc -1), a face mostly outside the frame (ROI hangs over the
    border: taps outside read 0), a face so large that its source footprint exceeds the staged
    tile (global-memory taps), and a single-frame clip.  Fewer frames than SMs in total."""
    H = W = 224
    specs = [("normal", 9), ("nodet", 4), ("offframe", 5), ("huge", 6), ("single", 1)]
    frames, lms, valids, lens = [], [], [], []
    for i, (kind, T) in enumerate(specs):
        f, lm, v = synth.video_clip(T, H, W, seed=300 + i, invalid_frac=0.0)
        if kind == "nodet":
            v[:] = 0
        elif kind == "offframe":
            lm = lm + np.array([150.0, -120.0])
        elif kind == "huge":
            c = lm.mean(axis=(0, 1), keepdims=True)
            lm = np.rint((lm - c) * 2.4 + c)              # inverse scale ~1.8: footprint ~165 x 176 px
        frames.append(f); lms.append(lm); valids.append(v); lens.append(T)
    F = np.concatenate(frames); LM = np.concatenate(lms); V = np.concatenate(valids)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for want_u8 in (False, True):
        res = L.lip_roi_batch(torch.from_numpy(F).cuda(), torch.from_numpy(off).cuda(), torch.from_numpy(LM).cuda(),
                              torch.from_numpy(V).cuda(), want_gray=True, want_u8=want_u8, want_f32=True, want_meta=True)
        gray_ref = O.bgr2gray(F)
        np.testing.assert_array_equal(res.gray.cpu().numpy(), gray_ref)
        mf = A.mean_face_landmarks()
        for i, (kind, T) in enumerate(specs):
            lo, hi = off[i], off[i + 1]
            got_f32 = res.lip_f32[lo:hi].cpu().numpy()
            if kind == "nodet":
                assert (res.crop_rc[lo:hi].cpu().numpy() == -1).all()
                np.testing.assert_array_equal(got_f32, O.video_feats_from_u8(np.zeros((T, 96, 96), np.uint8))[..., 0])
                continue
            ref, tf, org = O.extract_lip_frames_from_arrays(gray_ref[lo:hi], _as_list(lms[i], valids[i]), mf)
            np.testing.assert_array_equal(res.crop_rc[lo:hi].cpu().numpy(), org)
            feats_ref = O.video_feats_from_u8(ref)[..., 0]
            assert np.abs(got_f32 - feats_ref).max() <= (1.0 / 255.0) / 0.165 + 1e-6, kind
            assert (got_f32 != feats_ref).mean() < 2e-3, kind
            if want_u8:
                d = np.abs(res.lip_u8[lo:hi].cpu().numpy().astype(int) - ref.astype(int))
                assert d.max() <= 1 and (d != 0).mean() < 2e-3, kind


@pytest.mark.parametrize("H,W,T", [(288, 352, 40), (120, 176, 33), (96, 128, 150), (240, 320, 149), (224, 224, 297),
                                   (480, 640, 21)])      # 480 x 640: footprints larger than either kernel's tile
def test_frame_kernel_equals_generic_path(H, W, T):
    """The frame-owner kernel (gray wanted) and the generic work-queue path (no gray) are two
    implementations of the same arithmetic: identical ROI bytes, crop origins and transforms on
    frame sizes with partial last chunks, more / fewer frames than SMs, several clips."""
    cuts = sorted({0, T // 3, T // 3 + 1, T})
    lens = np.diff(cuts)
    clips = [synth.video_clip(int(n), H, W, seed=700 + i, invalid_frac=0.15) for i, n in enumerate(lens)]
    F = torch.from_numpy(np.concatenate([c[0] for c in clips])).cuda()
    LM = torch.from_numpy(np.concatenate([c[1] for c in clips])).cuda()
    V = torch.from_numpy(np.concatenate([c[2] for c in clips])).cuda()
    off = torch.tensor(cuts, dtype=torch.int64).cuda()
    for want_u8 in (False, True):
        a = L.lip_roi_batch(F, off, LM, V, want_gray=True, want_u8=want_u8, want_meta=True)
        b = L.lip_roi_batch(F, off, LM, V, want_gray=False, want_u8=want_u8, want_meta=True)
        assert torch.equal(a.lip_f32, b.lip_f32)
        assert torch.equal(a.crop_rc, b.crop_rc) and torch.equal(a.tforms, b.tforms)
        if want_u8:
            assert torch.equal(a.lip_u8, b.lip_u8)
    np.testing.assert_array_equal(a.gray.cpu().numpy(), O.bgr2gray(F.cpu().numpy()))
