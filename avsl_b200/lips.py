"""Lip-ROI front-end: host-side mirror of ``utils/lips_cropping.py``, the frame loop of
``preprocess/video_process.py::extract_lip_frames`` and the crop/normalise step of
``utils/hf_video_utils.py`` / ``utils/data_loading.py``.

Names, argument meaning and error behaviour follow the reference; the arithmetic runs in
libavfe.so on the GPU.  Face / landmark detection (dlib) and video decoding stay with the
caller: every function here starts from decoded frames and 68-point landmarks.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib

STABLE_IDS = (33, 36, 39, 42, 45)
STD_SIZE = (300, 300)
WINDOW_MARGIN = 12
IMAGE_CROP_SIZE = 88
IMAGE_MEAN = 0.421
IMAGE_STD = 0.165
_MEAN_FACE_PATH = Path(__file__).resolve().parent / "resources" / "mean_face_68.npy"
_MEAN_FACE_CACHE: dict = {}


def mean_face_landmarks() -> np.ndarray:
    """The 68x2 float64 reference face (reference: resources/20words_mean_face.npy)."""
    return np.load(_MEAN_FACE_PATH)


def _mean_face_dev(device, mean_face=None) -> torch.Tensor:
    if mean_face is not None:
        return torch.as_tensor(np.asarray(mean_face, dtype=np.float64)).contiguous().to(device)
    key = str(device)
    if key not in _MEAN_FACE_CACHE:
        _MEAN_FACE_CACHE[key] = torch.from_numpy(mean_face_landmarks()).contiguous().to(device)
    return _MEAN_FACE_CACHE[key]


def _dev() -> torch.device:
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(x, dtype, device=None) -> torch.Tensor:
    if torch.is_tensor(x):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if t.dtype != dtype:
        t = t.to(dtype)
    t = t.contiguous()
    if not t.is_cuda:
        t = t.pin_memory().to(device or _dev(), non_blocking=True)
    return t


# --------------------------------------------------------------------------- transform object
class SimilarityTransform:
    """What ``warp_img`` returns as ``tform``: ``.params`` (3x3 float64), callable on [N,2]
    points (``trans(cur_landmarks)``, preprocess/video_process.py:437) and ``.inverse``."""

    def __init__(self, matrix: np.ndarray, inverse_matrix: Optional[np.ndarray] = None):
        self.params = np.asarray(matrix, dtype=np.float64).reshape(3, 3)
        self._inv = None if inverse_matrix is None else np.asarray(inverse_matrix, dtype=np.float64).reshape(3, 3)

    @property
    def inverse(self) -> "SimilarityTransform":
        inv = self._inv if self._inv is not None else np.linalg.inv(self.params)
        return SimilarityTransform(inv, self.params)

    def __call__(self, coords) -> np.ndarray:
        c = np.array(coords, dtype=np.float64, ndmin=2)
        src = np.concatenate([c, np.ones((c.shape[0], 1))], axis=1)
        dst = src @ self.params.T
        dst[dst[:, 2] == 0, 2] = np.finfo(float).eps
        dst[:, :2] /= dst[:, 2:3]
        return dst[:, :2]


# --------------------------------------------------------------------------- V2
def landmarks_interpolate(landmarks: list) -> Optional[list]:
    """``utils/lips_cropping.py:60-89``: fill ``None`` entries by linear interpolation between
    neighbouring detections, replicate at both ends; returns ``None`` if nothing was detected."""
    T = len(landmarks)
    valid = np.array([lm is not None for lm in landmarks], dtype=np.uint8)
    if T == 0 or not valid.any():
        return None
    if valid.all():
        return list(landmarks)
    dense = np.zeros((T, 68, 2), dtype=np.float64)
    for i, lm in enumerate(landmarks):
        if lm is not None:
            dense[i] = np.asarray(lm, dtype=np.float64)
    dev = _dev()
    d_lm = _to_dev(dense, torch.float64, dev)
    d_valid = _to_dev(valid, torch.uint8, dev)
    d_off = torch.tensor([0, T], dtype=torch.int64, device=dev)
    out = torch.empty_like(d_lm)
    _lib.call("avfe_landmarks_interpolate", _lib.ptr(d_lm), _lib.ptr(d_valid), _lib.ptr(d_off), 1, T,
              _lib.ptr(out), _lib.stream_ptr())
    filled = out.cpu().numpy()
    # detected frames keep the caller's own arrays (the reference leaves them untouched)
    return [landmarks[i] if valid[i] else filled[i] for i in range(T)]


# --------------------------------------------------------------------------- V4 / V5
def _fit(src, dst, dev) -> Tuple[np.ndarray, np.ndarray]:
    s = _to_dev(np.asarray(src, dtype=np.float64), torch.float64, dev)
    d = _to_dev(np.asarray(dst, dtype=np.float64), torch.float64, dev)
    if s.shape != d.shape or s.dim() != 2 or s.shape[1] != 2:
        raise ValueError("src and dst must both be [n, 2]")
    out = torch.empty(18, dtype=torch.float64, device=dev)
    _lib.call("avfe_similarity_fit", _lib.ptr(s), _lib.ptr(d), int(s.shape[0]), _lib.ptr(out),
              _lib.stream_ptr())
    o = out.cpu().numpy()
    return o[:9].reshape(3, 3), o[9:].reshape(3, 3)


def _warp(img, inv_matrix: np.ndarray, std_size, dev) -> np.ndarray:
    g = _to_dev(img, torch.uint8, dev)
    if g.dim() != 2:
        raise ValueError("warp expects a 2-D grayscale uint8 image")
    m = _to_dev(np.asarray(inv_matrix, dtype=np.float64).reshape(9), torch.float64, dev)
    oh, ow = int(std_size[0]), int(std_size[1])
    out = torch.empty((oh, ow), dtype=torch.uint8, device=dev)
    _lib.call("avfe_warp_affine_u8", _lib.ptr(g), int(g.shape[0]), int(g.shape[1]), _lib.ptr(m),
              oh, ow, _lib.ptr(out), _lib.stream_ptr())
    return out.cpu().numpy()


def warp_img(src, dst, img, std_size=STD_SIZE):
    """``utils/lips_cropping.py:91-108``: similarity fit src->dst, warp ``img`` (2-D uint8) to
    ``std_size``; returns ``(warped uint8, tform)``."""
    dev = _dev()
    fwd, inv = _fit(src, dst, dev)
    tform = SimilarityTransform(fwd, inv)
    return _warp(img, inv, std_size, dev), tform


def apply_transform(transform, img, std_size=STD_SIZE):
    """``utils/lips_cropping.py:110-125``: warp with a given transform."""
    return _warp(img, transform.inverse.params, std_size, _dev())


# --------------------------------------------------------------------------- V7
def cut_patch(img, landmarks, height, width, threshold=5):
    """``utils/lips_cropping.py:127-163``: patch of half-size (height, width) centred on the mean
    landmark, clamped to the image."""
    dev = _dev()
    g = _to_dev(img, torch.uint8, dev)
    if g.dim() != 2:
        raise ValueError("cut_patch expects a 2-D uint8 image")
    lm = _to_dev(np.asarray(landmarks, dtype=np.float64), torch.float64, dev)
    hh, hw = int(round(height)), int(round(width))
    H, W = int(g.shape[0]), int(g.shape[1])
    if 2 * hh > H or 2 * hw > W:
        raise Exception("too much bias in height" if 2 * hh > H else "too much bias in width")
    out = torch.empty((2 * hh, 2 * hw), dtype=torch.uint8, device=dev)
    _lib.call("avfe_cut_patch_u8", _lib.ptr(g), H, W, _lib.ptr(lm), int(lm.shape[0]), hh, hw,
              _lib.ptr(out), None, _lib.stream_ptr())
    return out.cpu().numpy()


# --------------------------------------------------------------------------- fused batch op
class LipBatch:
    """Device-resident result of :func:`lip_roi_batch` (tensors are ``None`` when not asked for)."""

    __slots__ = ("gray", "lip_u8", "lip_f32", "crop_rc", "tforms", "clip_offsets")

    def __init__(self, gray, lip_u8, lip_f32, crop_rc, tforms, clip_offsets):
        self.gray, self.lip_u8, self.lip_f32 = gray, lip_u8, lip_f32
        self.crop_rc, self.tforms, self.clip_offsets = crop_rc, tforms, clip_offsets


def lip_workspace_bytes(n_frames: int) -> int:
    """Bytes of scratch ``avfe_lip_roi_batch`` / ``avfe_lip_roi_collate`` need for ``n_frames``."""
    return int(_lib.load().avfe_lip_workspace_bytes(int(n_frames)))


def _workspace(workspace: Optional[torch.Tensor], N: int, dev) -> torch.Tensor:
    need = lip_workspace_bytes(N)
    if workspace is None:
        return torch.empty(need, dtype=torch.uint8, device=dev)
    if not (workspace.is_cuda and workspace.dtype == torch.uint8 and workspace.is_contiguous()
            and workspace.numel() >= need and workspace.data_ptr() % 16 == 0):
        raise ValueError(f"workspace must be a contiguous 16-byte aligned CUDA uint8 tensor of >= {need} bytes")
    return workspace


def lip_roi_batch(frames: torch.Tensor, clip_offsets: torch.Tensor, landmarks: torch.Tensor,
                  lm_valid: Optional[torch.Tensor] = None, *, mean_face=None,
                  tforms_in: Optional[torch.Tensor] = None, want_gray: bool = True,
                  want_u8: bool = False, want_f32: bool = True, want_meta: bool = False,
                  roi: int = 96, crop: int = IMAGE_CROP_SIZE, std_size: int = 300,
                  window: int = WINDOW_MARGIN, image_mean: float = IMAGE_MEAN,
                  image_std: float = IMAGE_STD, out: Optional[LipBatch] = None,
                  workspace: Optional[torch.Tensor] = None) -> LipBatch:
    """One fused pass over a batch of clips stored back to back (all tensors on the GPU).

    frames [N,H,W,3] BGR uint8 (or [N,H,W] gray), clip_offsets int64 [n_clips+1], landmarks
    float64 [N,68,2], lm_valid uint8 [N] (0 = detection failed).  See ``avfe_lip_roi_batch`` in
    include/avfe.h.  ``out`` reuses previously returned buffers and ``workspace`` a caller-owned
    scratch tensor of ``lip_workspace_bytes(N)`` bytes (steady-state loops, CUDA-graph capture:
    nothing is allocated then); a workspace must not be shared by calls that may overlap."""
    _lib.require_cuda()
    host_frames = not frames.is_cuda
    if host_frames:
        # Zero-copy: page-locked host memory is mapped into the device's address space (unified
        # addressing), so the kernel can pull just the ROI footprints across PCIe instead of the
        # whole frames being copied first.  Only without gray output (that needs every pixel).
        if not frames.is_pinned():
            raise ValueError("lip_roi_batch takes CUDA tensors or PINNED host frames; use extract_lip_frames for host arrays")
        if want_gray:
            raise ValueError("host-resident (zero-copy) frames cannot be combined with want_gray=True")
    if frames.dtype != torch.uint8 or not frames.is_contiguous():
        raise ValueError("frames must be contiguous uint8")
    if frames.dim() == 4 and frames.shape[-1] == 3:
        channels = 3
    elif frames.dim() == 3:
        channels = 1
    else:
        raise ValueError("frames must be [N,H,W,3] (BGR) or [N,H,W] (gray)")
    N, H, W = int(frames.shape[0]), int(frames.shape[1]), int(frames.shape[2])
    dev = clip_offsets.device if host_frames else frames.device
    n_clips = int(clip_offsets.numel()) - 1
    if clip_offsets.dtype != torch.int64 or not clip_offsets.is_cuda:
        raise ValueError("clip_offsets must be a CUDA int64 tensor")
    if landmarks.dtype != torch.float64 or tuple(landmarks.shape) != (N, 68, 2) or not landmarks.is_contiguous():
        raise ValueError("landmarks must be contiguous float64 [N,68,2]")
    mf = _mean_face_dev(dev, mean_face)
    want_gray = want_gray and channels == 3
    if out is None:
        gray = torch.empty((N, H, W), dtype=torch.uint8, device=dev) if want_gray else None
        lip_u8 = torch.empty((N, roi, roi), dtype=torch.uint8, device=dev) if want_u8 else None
        lip_f32 = torch.empty((N, crop, crop), dtype=torch.float32, device=dev) if want_f32 else None
        crop_rc = torch.empty((N, 2), dtype=torch.int32, device=dev) if want_meta else None
        tforms = torch.empty((N, 18), dtype=torch.float64, device=dev) if want_meta else None
        out = LipBatch(gray, lip_u8, lip_f32, crop_rc, tforms, clip_offsets)
    with torch.cuda.device(dev):
        ws = _workspace(workspace, N, dev)
        ws_bytes = int(ws.numel())
        _lib.call("avfe_lip_roi_batch", _lib.ptr(frames), channels, N, H, W, _lib.ptr(clip_offsets),
                  n_clips, _lib.ptr(landmarks), _lib.ptr(lm_valid), _lib.ptr(mf), _lib.ptr(tforms_in),
                  std_size, roi, crop, window, float(image_mean), float(image_std),
                  _lib.ptr(out.gray), _lib.ptr(out.lip_u8), _lib.ptr(out.lip_f32),
                  _lib.ptr(out.crop_rc), _lib.ptr(out.tforms), _lib.ptr(ws), ws_bytes,
                  _lib.stream_ptr())
    return out


def video_frames_for_audio(n_audio_samples: int, sample_rate: int = 16000, fps: int = 25) -> int:
    """Frames the reference keeps for an audio of ``n_audio_samples``
    (avsl/whisper_flamingo_ft_ami.py:299): Python ``round`` of the duration times 25."""
    return round(n_audio_samples / sample_rate * fps)


def lip_roi_collate(frames: torch.Tensor, clip_offsets: torch.Tensor, landmarks: torch.Tensor,
                    lm_valid: Optional[torch.Tensor] = None, *, T_pad: int,
                    keep_frames: Optional[torch.Tensor] = None, mean_face=None,
                    tforms_in: Optional[torch.Tensor] = None, want_gray: bool = True,
                    roi: int = 96, crop: int = IMAGE_CROP_SIZE, std_size: int = 300,
                    window: int = WINDOW_MARGIN, image_mean: float = IMAGE_MEAN,
                    image_std: float = IMAGE_STD, out: Optional[Dict[str, torch.Tensor]] = None,
                    workspace: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """:func:`lip_roi_batch` writing straight into the padded batch the encoder consumes
    (``features, x_v = self.model.encoder(input_ids, video, ..., padding_mask=padding_mask)``,
    avsl/whisper_flamingo_ft_ami.py:527): the per-sample trim of ``__getitem__`` (:299-302) and
    the zero padding + mask of the upstream ``WhisperVideoCollatorWithPadding`` (:126, :686) are
    folded into the kernel's output addressing, so the packed ``[N,88,88]`` tensor is never built.

    keep_frames int64 [n_clips] (CUDA) = frames kept per clip after the trim (None: all);
    T_pad = frames per clip in the batch (the collator uses the longest kept clip).
    Returns ``video`` float32 [B,1,T_pad,88,88], ``padding_mask`` bool [B,T_pad] (True = padding)
    and ``gray`` uint8 [N,H,W] (if wanted)."""
    _lib.require_cuda()
    if not frames.is_cuda or frames.dtype != torch.uint8 or not frames.is_contiguous():
        raise ValueError("frames must be a contiguous CUDA uint8 tensor")
    if frames.dim() == 4 and frames.shape[-1] == 3:
        channels = 3
    elif frames.dim() == 3:
        channels = 1
    else:
        raise ValueError("frames must be [N,H,W,3] (BGR) or [N,H,W] (gray)")
    if T_pad < 1:
        raise ValueError("T_pad must be >= 1")
    N, H, W = int(frames.shape[0]), int(frames.shape[1]), int(frames.shape[2])
    dev = frames.device
    n_clips = int(clip_offsets.numel()) - 1
    if clip_offsets.dtype != torch.int64 or not clip_offsets.is_cuda:
        raise ValueError("clip_offsets must be a CUDA int64 tensor")
    if landmarks.dtype != torch.float64 or tuple(landmarks.shape) != (N, 68, 2) or not landmarks.is_contiguous():
        raise ValueError("landmarks must be contiguous float64 [N,68,2]")
    if keep_frames is not None and (keep_frames.dtype != torch.int64 or not keep_frames.is_cuda or
                                    keep_frames.numel() != n_clips):
        raise ValueError("keep_frames must be a CUDA int64 tensor [n_clips]")
    mf = _mean_face_dev(dev, mean_face)
    want_gray = want_gray and channels == 3
    if out is None:
        out = {"video": torch.empty((n_clips, 1, T_pad, crop, crop), dtype=torch.float32, device=dev),
               "padding_mask_u8": torch.empty((n_clips, T_pad), dtype=torch.uint8, device=dev)}
        if want_gray:
            out["gray"] = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(workspace, N, dev)
        ws_bytes = int(ws.numel())
        _lib.call("avfe_lip_roi_collate", _lib.ptr(frames), channels, N, H, W, _lib.ptr(clip_offsets),
                  n_clips, _lib.ptr(landmarks), _lib.ptr(lm_valid), _lib.ptr(mf), _lib.ptr(tforms_in),
                  std_size, roi, crop, window, float(image_mean), float(image_std),
                  _lib.ptr(keep_frames), int(T_pad), _lib.ptr(out.get("gray")), _lib.ptr(out["video"]),
                  _lib.ptr(out["padding_mask_u8"]), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
    out["padding_mask"] = out["padding_mask_u8"].view(torch.bool)
    return out


def _landmark_arrays(landmarks: Sequence) -> Tuple[np.ndarray, np.ndarray]:
    T = len(landmarks)
    dense = np.zeros((T, 68, 2), dtype=np.float64)
    valid = np.zeros(T, dtype=np.uint8)
    for i, lm in enumerate(landmarks):
        if lm is not None:
            dense[i] = np.asarray(lm, dtype=np.float64)
            valid[i] = 1
    return dense, valid


# --------------------------------------------------------------------------- V3..V7 driver
def extract_lip_frames(frames: np.ndarray, landmarks: Sequence, mean_face_path=None,
                       width_roi: int = 96, height_roi: int = 96, start_idx: int = 48,
                       stop_idx: int = 68, to_grayscale: bool = True,
                       max_frames: Optional[int] = None) -> np.ndarray:
    """The frame loop of ``extract_lip_frames`` (preprocess/video_process.py:305-490) for decoded
    frames: ``frames`` is [T,H,W,3] BGR (as ``cv2.VideoCapture`` yields) or [T,H,W] gray uint8,
    ``landmarks`` a list of T detections ([68,2] or ``None``).  Returns uint8 [T,96,96], or an
    empty array on failure exactly like the reference (which swallows every error)."""
    try:
        if (start_idx, stop_idx) != (48, 68) or width_roi != height_roi or not to_grayscale:
            raise NotImplementedError("only the reference defaults (mouth 48:68, square gray ROI) are built")
        frames = np.asarray(frames)
        if frames.size == 0:
            print("No frames loaded, cannot extract lips.")
            return np.array([])
        if max_frames is not None and len(frames) > max_frames:
            stride = max(1, len(frames) // max_frames)            # load_video's frame_stride
            frames = frames[::stride][:max_frames]
            landmarks = list(landmarks)[::stride][:max_frames]
        if len(landmarks) != len(frames):
            raise ValueError("one landmark entry per frame is required")
        dense, valid = _landmark_arrays(landmarks)
        if not valid.any():
            print("No face detected or landmarks couldn't be interpolated")
            return np.array([])
        dev = _dev()
        mean_face = None if mean_face_path is None else np.load(mean_face_path)
        d_frames = _to_dev(frames, torch.uint8, dev)
        T = int(d_frames.shape[0])
        res = lip_roi_batch(d_frames, torch.tensor([0, T], dtype=torch.int64, device=dev),
                            _to_dev(dense, torch.float64, dev), _to_dev(valid, torch.uint8, dev),
                            mean_face=mean_face, want_gray=False, want_u8=True, want_f32=False,
                            roi=width_roi)
        return res.lip_u8.cpu().numpy()
    except Exception as e:  # reference: print and return an empty array
        print(f"Critical Error extracting lip frames: {str(e)}")
        return np.array([])


# --------------------------------------------------------------------------- V8
def load_video_feats(frames, train: bool = False, image_crop_size: int = IMAGE_CROP_SIZE,
                     image_mean: float = IMAGE_MEAN, image_std: float = IMAGE_STD):
    """The arithmetic of ``load_video_feats_from_decord_reader`` (utils/hf_video_utils.py:103-138)
    on the decoded uint8 frames, exactly as its call site runs it
    (``safe_load_video_feats_from_hf_object``, avsl/whisper_flamingo_ft_ami.py:279-286):

    * ``[T,H,W,3]`` (what decord returns, RGB): float64 ``np.dot`` with ``[0.2989, 0.5870, 0.1140]``,
      then float32 ``/255`` if the stack's maximum exceeds 1.0 (:105, :116-117);
    * ``[T,H,W]`` / ``[T,H,W,1]``: float32 ``/255`` (:107-108, :114-115; also
      ``load_video_features``, utils/data_loading.py:101-118);
    * centre crop, or ``cv2.resize`` to (crop, crop) when a frame side is smaller than the crop
      (:120-132); ``(x - mean) / std``; trailing channel axis; float32 -> ``[T,crop,crop,1]``.

    numpy in -> numpy out; CUDA tensor in -> CUDA tensor out."""
    is_t = torch.is_tensor(frames)
    dev = frames.device if (is_t and frames.is_cuda) else _dev()
    t = _to_dev(frames, torch.uint8, dev)
    channels = 1
    if t.dim() == 4 and t.shape[-1] == 3:
        channels = 3
    elif t.dim() == 4 and t.shape[-1] == 1:
        t = t.squeeze(-1).contiguous()
    if t.dim() != 3 + (channels == 3):
        raise ValueError(f"Expected 3D frames array after processing, got shape: {tuple(t.shape)}")
    N, H, W = (int(s) for s in t.shape[:3])
    out = torch.empty((N, image_crop_size, image_crop_size, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = torch.empty(int(_lib.load().avfe_video_feats_workspace_bytes()), dtype=torch.uint8, device=dev)
        _lib.call("avfe_video_feats", _lib.ptr(t), channels, N, H, W, int(image_crop_size), float(image_mean),
                  float(image_std), _lib.ptr(out), _lib.ptr(ws), int(ws.numel()), _lib.stream_ptr())
    if is_t and frames.is_cuda:
        return out
    return out.cpu().numpy()


def bgr2gray(frames):
    """``cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)`` for [..., H, W, 3] uint8 (bit-exact)."""
    is_t = torch.is_tensor(frames)
    dev = frames.device if (is_t and frames.is_cuda) else _dev()
    t = _to_dev(frames, torch.uint8, dev)
    if t.shape[-1] != 3:
        raise ValueError("expected [..., H, W, 3] BGR input")
    H, W = int(t.shape[-3]), int(t.shape[-2])
    N = int(t.numel() // (H * W * 3)) if H * W else 0
    out = torch.empty(t.shape[:-1], dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call("avfe_bgr2gray_u8", _lib.ptr(t), N, H, W, _lib.ptr(out), _lib.stream_ptr())
    if is_t and frames.is_cuda:
        return out
    return out.cpu().numpy()


def trim_video_to_audio(video_feats, n_audio_samples: int, sample_rate: int = 16000):
    """``avsl/whisper_flamingo_ft_ami.py:299-302``: keep at most round(samples/sr*25) frames."""
    max_len = round(n_audio_samples / sample_rate * 25)
    return video_feats[:max_len] if len(video_feats) > max_len else video_feats
