"""Oracle: landmark-driven lip-ROI path (CPU, numpy float64).  TEST INFRASTRUCTURE ONLY.

Restates, in the reference's own operation order:

* ``cv2.cvtColor(BGR2GRAY)``                    preprocess/video_process.py:201-214
* ``linear_interpolate`` / ``landmarks_interpolate``  utils/lips_cropping.py:41-89
* ``warp_img`` / ``apply_transform``            utils/lips_cropping.py:91-125
* ``cut_patch``                                 utils/lips_cropping.py:127-163
* the window loop of ``extract_lip_frames``     preprocess/video_process.py:369-475
* crop 88 + normalise                           utils/hf_video_utils.py:103-138,
                                                utils/data_loading.py:46-66,101-118

PARITY UNPINNED for the similarity fit and the warp: their arithmetic lives in scikit-image
(``requirements.txt:15``, unpinned), which cannot be installed in this image.  The code below
follows skimage's published algorithm (``transform._geometric._umeyama``,
``SimilarityTransform.inverse``/``__call__``, ``transform._warps.warp`` ->
``_warps_cy._warp_fast`` order 1, mode 'constant', cval 0, ``img_as_float`` input,
``_clip_warp_output`` in its >=0.20 form) and is cross-checked in ``tests/`` against
``scipy.ndimage.map_coordinates`` and ``cv2.warpAffine``.  The BGR->gray formula is pinned
bit-exact against ``cv2`` and the crop/normalise (uint8, RGB and small-frame ``cv2.resize``
branches) against the reference function itself.
"""
from __future__ import annotations

from collections import deque
from typing import List, Optional, Sequence, Tuple

import numpy as np

STABLE_IDS = [33, 36, 39, 42, 45]     # preprocess/video_process.py:398
STD_SIZE = (300, 300)                 # preprocess/video_process.py:399
WINDOW_MARGIN = 12                    # preprocess/video_process.py:370
MOUTH_START, MOUTH_STOP = 48, 68      # preprocess/video_process.py:311-312
ROI = 96                              # width_roi / height_roi, preprocess/video_process.py:309-310
CROP = 88                             # avsl/whisper_flamingo_ft_ami.py:282
IMAGE_MEAN, IMAGE_STD = 0.421, 0.165  # avsl/whisper_flamingo_ft_ami.py:283-284


# ----------------------------------------------------------------------------- V1
def bgr2gray(bgr: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(frame, COLOR_BGR2GRAY) for uint8 (preprocess/video_process.py:214):
    15-bit fixed point, Y = (3735*B + 19235*G + 9798*R + 16384) >> 15."""
    b = bgr[..., 0].astype(np.uint32)
    g = bgr[..., 1].astype(np.uint32)
    r = bgr[..., 2].astype(np.uint32)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


# ----------------------------------------------------------------------------- V2
def linear_interpolate(landmarks: list, start_idx: int, stop_idx: int) -> list:
    """utils/lips_cropping.py:41-58."""
    start_landmarks = landmarks[start_idx]
    stop_landmarks = landmarks[stop_idx]
    delta = stop_landmarks - start_landmarks
    for idx in range(1, stop_idx - start_idx):
        landmarks[start_idx + idx] = start_landmarks + idx / float(stop_idx - start_idx) * delta
    return landmarks


def landmarks_interpolate(landmarks: list) -> Optional[list]:
    """utils/lips_cropping.py:60-89.  ``None`` entries are filled by linear interpolation
    between the neighbouring detections and by replication at both ends; all-``None`` -> None."""
    landmarks = list(landmarks)
    valid = [i for i, lm in enumerate(landmarks) if lm is not None]
    if not valid:
        return None
    for k in range(1, len(valid)):
        if valid[k] - valid[k - 1] == 1:
            continue
        landmarks = linear_interpolate(landmarks, valid[k - 1], valid[k])
    valid = [i for i, lm in enumerate(landmarks) if lm is not None]
    if valid:
        for i in range(0, valid[0]):
            landmarks[i] = landmarks[valid[0]]
        for i in range(valid[-1] + 1, len(landmarks)):
            landmarks[i] = landmarks[valid[-1]]
    return landmarks


# ----------------------------------------------------------------------------- V4 (fit)
def umeyama(src: np.ndarray, dst: np.ndarray, estimate_scale: bool = True) -> np.ndarray:
    """skimage.transform._geometric._umeyama — what estimate_transform('similarity', src, dst)
    evaluates (utils/lips_cropping.py:104).  Returns the 3x3 homogeneous matrix."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    num, dim = src.shape
    src_mean = src.mean(axis=0)
    dst_mean = dst.mean(axis=0)
    src_demean = src - src_mean
    dst_demean = dst - dst_mean
    A = dst_demean.T @ src_demean / num
    d = np.ones((dim,), dtype=np.float64)
    if np.linalg.det(A) < 0:
        d[dim - 1] = -1
    T = np.eye(dim + 1, dtype=np.float64)
    U, S, V = np.linalg.svd(A)
    rank = np.linalg.matrix_rank(A)
    if rank == 0:
        return np.nan * T
    elif rank == dim - 1:
        if np.linalg.det(U) * np.linalg.det(V) > 0:
            T[:dim, :dim] = U @ V
        else:
            s = d[dim - 1]
            d[dim - 1] = -1
            T[:dim, :dim] = U @ np.diag(d) @ V
            d[dim - 1] = s
    else:
        T[:dim, :dim] = U @ np.diag(d) @ V
    if estimate_scale:
        scale = 1.0 / src_demean.var(axis=0).sum() * (S @ d)
    else:
        scale = 1.0
    T[:dim, dim] = dst_mean - scale * (T[:dim, :dim] @ src_mean.T)
    T[:dim, :dim] *= scale
    return T


class SimilarityTransform:
    """Minimal stand-in for skimage's transform object returned by ``warp_img``: holds
    ``params`` (3x3), is callable on [N,2] points, and exposes ``inverse``."""

    def __init__(self, matrix: np.ndarray):
        self.params = np.asarray(matrix, dtype=np.float64)

    @property
    def inverse(self) -> "SimilarityTransform":
        return SimilarityTransform(np.linalg.inv(self.params))

    def __call__(self, coords: np.ndarray) -> np.ndarray:
        # skimage ProjectiveTransform._apply_mat
        coords = np.array(coords, copy=True, ndmin=2, dtype=np.float64)
        src = np.concatenate([coords, np.ones((coords.shape[0], 1))], axis=1)
        dst = src @ self.params.T
        dst[dst[:, 2] == 0, 2] = np.finfo(float).eps
        dst[:, :2] /= dst[:, 2:3]
        return dst[:, :2]


# ----------------------------------------------------------------------------- V4/V5 (warp)
def warp_float(img_u8: np.ndarray, inv_matrix: np.ndarray, output_shape=STD_SIZE,
               rows: Optional[Tuple[int, int]] = None,
               cols: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """skimage.transform.warp(img, inverse_map=tform.inverse, output_shape) for a 2-D uint8
    image: float64 result in [0,1].  ``inv_matrix`` = np.linalg.inv(tform.params) — skimage
    hands ``tform.inverse.params`` to ``_warp_fast``.  ``rows``/``cols`` restrict evaluation to
    a window of the output (identical values; used to skip the 90% of the 300x300 frame that
    ``cut_patch`` throws away)."""
    img = np.ascontiguousarray(img_u8)
    assert img.ndim == 2 and img.dtype == np.uint8
    image = img.astype(np.float64) / 255.0                      # img_as_float
    H, W = image.shape
    M = np.asarray(inv_matrix, dtype=np.float64)
    r0, r1 = rows if rows is not None else (0, output_shape[0])
    c0, c1 = cols if cols is not None else (0, output_shape[1])
    tfr = np.arange(r0, r1, dtype=np.float64)[:, None]
    tfc = np.arange(c0, c1, dtype=np.float64)[None, :]
    # _transform_affine(x=tfc, y=tfr): x_ = M0*x + M1*y + M2 ; y_ = M3*x + M4*y + M5
    c = M[0, 0] * tfc + M[0, 1] * tfr + M[0, 2]
    r = M[1, 0] * tfc + M[1, 1] * tfr + M[1, 2]
    if not (M[2, 0] == 0.0 and M[2, 1] == 0.0 and M[2, 2] == 1.0):   # _transform_projective
        z = M[2, 0] * tfc + M[2, 1] * tfr + M[2, 2]
        c = c / z
        r = r / z
    # bilinear_interpolation (skimage/_shared/interpolation.pxd), mode 'C', cval 0
    fr, fc = np.floor(r), np.floor(c)
    minr, minc = fr.astype(np.int64), fc.astype(np.int64)
    maxr, maxc = np.ceil(r).astype(np.int64), np.ceil(c).astype(np.int64)
    dr, dc = r - fr, c - fc

    def px(rr, cc):
        ok = (rr >= 0) & (rr < H) & (cc >= 0) & (cc < W)
        return np.where(ok, image[np.clip(rr, 0, H - 1), np.clip(cc, 0, W - 1)], 0.0)

    top = (1 - dc) * px(minr, minc) + dc * px(minr, maxc)
    bottom = (1 - dc) * px(maxr, minc) + dc * px(maxr, maxc)
    out = (1 - dr) * top + dr * bottom
    # _clip_warp_output (clip=True, order 1, mode 'constant', cval 0): range expanded to
    # include cval, i.e. [min(min,0), max(max,0)].
    lo = min(float(image.min()), 0.0) if image.size else 0.0
    hi = max(float(image.max()), 0.0) if image.size else 0.0
    np.clip(out, lo, hi, out=out)
    return out


def to_u8(warped: np.ndarray) -> np.ndarray:
    """``warped = warped * 255; warped.astype('uint8')`` (utils/lips_cropping.py:106-107)."""
    return (warped * 255).astype("uint8")


def warp_img(src, dst, img, std_size=STD_SIZE):
    """utils/lips_cropping.py:91-108 -> (uint8 [std_size], tform)."""
    tform = SimilarityTransform(umeyama(src, dst, True))
    warped = warp_float(img, tform.inverse.params, std_size)
    return to_u8(warped), tform


def apply_transform(transform, img, std_size=STD_SIZE):
    """utils/lips_cropping.py:110-125."""
    return to_u8(warp_float(img, transform.inverse.params, std_size))


# ----------------------------------------------------------------------------- V7
def cut_patch_origin(landmarks: np.ndarray, height: int, width: int,
                     img_shape=STD_SIZE, threshold: int = 5) -> Tuple[int, int]:
    """Row/col of the top-left corner chosen by ``cut_patch`` (utils/lips_cropping.py:141-162)."""
    center_x, center_y = np.mean(landmarks, axis=0)
    if center_y - height < 0:
        center_y = height
    if center_y - height < 0 - threshold:
        raise Exception("too much bias in height")
    if center_x - width < 0:
        center_x = width
    if center_x - width < 0 - threshold:
        raise Exception("too much bias in width")
    if center_y + height > img_shape[0]:
        center_y = img_shape[0] - height
    if center_y + height > img_shape[0] + threshold:
        raise Exception("too much bias in height")
    if center_x + width > img_shape[1]:
        center_x = img_shape[1] - width
    if center_x + width > img_shape[1] + threshold:
        raise Exception("too much bias in width")
    return int(round(center_y) - round(height)), int(round(center_x) - round(width))


def cut_patch(img, landmarks, height, width, threshold=5):
    """utils/lips_cropping.py:127-163."""
    r0, c0 = cut_patch_origin(landmarks, height, width, img.shape, threshold)
    return np.copy(img[r0: r0 + 2 * int(round(height)), c0: c0 + 2 * int(round(width))])


# ----------------------------------------------------------------------------- V3..V7 driver
def extract_lip_frames_from_arrays(frames_gray: np.ndarray, landmarks: Sequence,
                                   mean_face: np.ndarray, width_roi: int = ROI,
                                   height_roi: int = ROI, start_idx: int = MOUTH_START,
                                   stop_idx: int = MOUTH_STOP, full_warp: bool = False):
    """The body of ``extract_lip_frames`` after frame loading and landmark detection
    (preprocess/video_process.py:392-485), fed with grayscale frames [T,H,W] uint8 and a
    list of T landmark arrays (``None`` = detection failed).

    Returns ``(rois u8 [T,h,w], tforms f64 [T,3,3], origins int [T,2])`` or
    ``(np.array([]), None, None)`` when no face was ever detected.
    ``full_warp=True`` evaluates the whole 300x300 warp like the reference does (slow);
    the default evaluates only the window ``cut_patch`` keeps — same values.
    """
    landmarks = landmarks_interpolate(list(landmarks))
    if landmarks is None:
        return np.array([]), None, None
    T = len(frames_gray)
    margin = min(T, WINDOW_MARGIN)
    seq, tforms, origins = [], [], []
    q_frame, q_landmarks = deque(), deque()
    trans = None

    def emit(tf: SimilarityTransform, frame, cur_landmarks):
        t_lm = tf(cur_landmarks)
        r0, c0 = cut_patch_origin(t_lm[start_idx:stop_idx], height_roi // 2, width_roi // 2, STD_SIZE)
        hh, ww = 2 * (height_roi // 2), 2 * (width_roi // 2)
        if full_warp:
            patch = to_u8(warp_float(frame, tf.inverse.params, STD_SIZE))[r0:r0 + hh, c0:c0 + ww].copy()
        else:
            patch = to_u8(warp_float(frame, tf.inverse.params, STD_SIZE, (r0, r0 + hh), (c0, c0 + ww)))
        seq.append(patch)
        tforms.append(tf.params.copy())
        origins.append((r0, c0))

    for frame_idx in range(T):
        q_landmarks.append(landmarks[frame_idx])
        q_frame.append(frames_gray[frame_idx])
        if len(q_frame) == margin:
            smoothed = np.mean(q_landmarks, axis=0)
            cur_landmarks = q_landmarks.popleft()
            cur_frame = q_frame.popleft()
            trans = SimilarityTransform(umeyama(smoothed[STABLE_IDS, :], mean_face[STABLE_IDS, :], True))
            emit(trans, cur_frame, cur_landmarks)
    while q_frame:
        cur_frame = q_frame.popleft()
        cur_landmarks = q_landmarks.popleft()
        if trans is None:
            continue
        emit(trans, cur_frame, cur_landmarks)
    if not seq:
        return np.array([]), None, None
    return np.array(seq), np.array(tforms), np.array(origins, dtype=np.int32)


# ----------------------------------------------------------------------------- V8
def video_feats_from_u8(frames_u8: np.ndarray, image_crop_size: int = CROP,
                        image_mean: float = IMAGE_MEAN, image_std: float = IMAGE_STD) -> np.ndarray:
    """utils/hf_video_utils.py:113-138 for uint8 [T,H,W] input, then ``.astype(np.float32)``
    (avsl/whisper_flamingo_ft_ami.py:286): /255 in float32, centre crop, (x-mean)/std, [...,None]."""
    frames = frames_u8.astype(np.float32) / 255.0
    H, W = frames.shape[1], frames.shape[2]
    sh, sw = (H - image_crop_size) // 2, (W - image_crop_size) // 2
    assert sh >= 0 and sw >= 0, "small-frame cv2.resize fallback is out of scope"
    frames = frames[:, sh:sh + image_crop_size, sw:sw + image_crop_size]
    frames = (frames - image_mean) / image_std
    return np.expand_dims(frames, axis=-1).astype(np.float32)


def resize_linear(src: np.ndarray, dsize: int) -> np.ndarray:
    """``cv2.resize(frame, (dsize, dsize))`` (INTER_LINEAR, utils/hf_video_utils.py:129) for a 2-D
    float32 / float64 frame, restating OpenCV's own generic code path (modules/imgproc/src/resize.cpp,
    OpenCV 4.x: coordinates ``fx = (float)((dx+0.5)*scale_x - 0.5)``, ``sx = floor(fx)``, the
    horizontal border rule ``sx < 0 -> (0, fx=0)``, ``sx >= W-1 -> (W-1, fx=0)``, row indices clipped
    to the frame, float32 coefficients, HResizeLinear then VResizeLinear with separately rounded
    products and sums).  Bit-exact against ``cv2.resize`` with IPP switched off
    (``cv2.ipp.setUseIPP(False)``); x86 wheels use IPP by default, whose float32 results differ
    from OpenCV's own code by a few 1e-6 (tests/test_oracle_lips.py pins both)."""
    H, W = src.shape
    T = src.dtype.type

    def coords(n_src):
        scale = 1.0 / (dsize / n_src)
        f = ((np.arange(dsize, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        return s, (f - s.astype(np.float32)).astype(np.float32)

    sx, fx = coords(W)
    fx = np.where((sx < 0) | (sx >= W - 1), np.float32(0), fx).astype(np.float32)
    sx = np.clip(sx, 0, W - 1)
    x1 = np.minimum(sx + 1, W - 1)
    sy, fy = coords(H)
    r0, r1 = np.clip(sy, 0, H - 1), np.clip(sy + 1, 0, H - 1)
    a0, a1 = (np.float32(1) - fx).astype(T), fx.astype(T)
    b0, b1 = (np.float32(1) - fy).astype(T), fy.astype(T)
    hrow = src[:, sx] * a0[None, :] + src[:, x1] * a1[None, :]
    return hrow[r0] * b0[:, None] + hrow[r1] * b1[:, None]


def video_feats_from_frames(frames: np.ndarray, image_crop_size: int = CROP, image_mean: float = IMAGE_MEAN,
                            image_std: float = IMAGE_STD) -> np.ndarray:
    """utils/hf_video_utils.py:103-138 for the uint8 frames a decord reader returns ([T,H,W,3] RGB,
    [T,H,W,1] or [T,H,W]), then ``.astype(np.float32)`` (avsl/whisper_flamingo_ft_ami.py:286).
    The RGB weights are applied channel by channel in float64 (numpy hands ``np.dot`` to BLAS; its
    summation order can move the last float64 bit, never the float32 the bright branch casts to --
    tests/test_oracle_lips.py walks all 2^24 triples)."""
    frames = np.asarray(frames)
    if frames.ndim == 4 and frames.shape[3] == 3:
        f64 = frames.astype(np.float64)
        frames = (f64[..., 0] * 0.2989 + f64[..., 1] * 0.5870) + f64[..., 2] * 0.1140      # :105
    elif frames.ndim == 4 and frames.shape[3] == 1:
        frames = frames.squeeze(axis=3)
    if frames.ndim != 3:
        raise ValueError(f"Expected 3D frames array after processing, got shape: {frames.shape}")
    if frames.dtype == np.uint8:                                                           # :114-117
        frames = frames.astype(np.float32) / 255.0
    elif frames.max() > 1.0:
        frames = frames.astype(np.float32) / 255.0
    H, W = frames.shape[1], frames.shape[2]
    start_h, start_w = (H - image_crop_size) // 2, (W - image_crop_size) // 2
    if start_h >= 0 and start_w >= 0:
        frames = frames[:, start_h:start_h + image_crop_size, start_w:start_w + image_crop_size]
    else:                                                                                  # :126-132
        frames = np.stack([resize_linear(f, image_crop_size) for f in frames])
    frames = (frames - image_mean) / image_std
    return np.expand_dims(frames, axis=-1).astype(np.float32)


def trim_video_to_audio(video_feats: np.ndarray, n_audio_samples: int, sample_rate: int = 16000):
    """avsl/whisper_flamingo_ft_ami.py:299-302."""
    max_len = round(n_audio_samples / sample_rate * 25)
    return video_feats[:max_len] if len(video_feats) > max_len else video_feats
