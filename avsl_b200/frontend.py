"""Batched AV front-end: what ``AmiVideoHFDataset.__getitem__`` computes per sample
(``avsl/whisper_flamingo_ft_ami.py:187-313``: pad_or_trim -> log-mel; lip frames -> crop 88 ->
normalise -> trim) and what ``extract_lip_frames`` computes per video, for a whole batch of
utterances in a handful of kernel launches on one GPU.

Utterances are independent, so multi-GPU operation is one process per GPU, each taking the
utterances ``i % world_size == rank`` (``shard``); there is no collective on the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .audio import (HOP_LENGTH, N_SAMPLES, SAMPLE_RATE, log_mel_spectrogram, log_mel_spectrogram_ragged,
                    mel_filters)
from .lips import (IMAGE_CROP_SIZE, IMAGE_MEAN, IMAGE_STD, LipBatch, lip_roi_batch, lip_workspace_bytes,
                   video_frames_for_audio)

FPS = 25


@dataclass
class PackedBatch:
    """A batch of utterances stored back to back (host-pinned or device tensors).

    audio          float32 [sum L_i]     waveforms, 16 kHz
    audio_offsets  int64   [U+1]
    frames         uint8   [N,H,W,3]     BGR frames of all clips, N = sum T_i
    clip_offsets   int64   [U+1]
    landmarks      float64 [N,68,2]
    lm_valid       uint8   [N]           0 = detection failed (reference: None)
    """
    audio: torch.Tensor
    audio_offsets: torch.Tensor
    frames: torch.Tensor
    clip_offsets: torch.Tensor
    landmarks: torch.Tensor
    lm_valid: Optional[torch.Tensor]

    FIELDS = ("audio", "audio_offsets", "frames", "clip_offsets", "landmarks", "lm_valid")

    @property
    def n_utts(self) -> int:
        return int(self.audio_offsets.numel()) - 1

    def nbytes(self) -> int:
        return sum(getattr(self, f).numel() * getattr(self, f).element_size()
                   for f in self.FIELDS if getattr(self, f) is not None)

    def pin(self) -> "PackedBatch":
        return PackedBatch(*[None if getattr(self, f) is None else getattr(self, f).cpu().contiguous().pin_memory()
                             for f in self.FIELDS])

    def to(self, device, non_blocking: bool = True, frames_stay_on_host: bool = False) -> "PackedBatch":
        """Copy to ``device``.  ``frames_stay_on_host`` leaves the (pinned) frames where they are:
        the zero-copy mode of the lip kernel reads them through the mapped host pointer."""
        return PackedBatch(*[None if getattr(self, f) is None else
                             (getattr(self, f) if (frames_stay_on_host and f == "frames") else
                              getattr(self, f).to(device, non_blocking=non_blocking))
                             for f in self.FIELDS])


def pack_utterances(audios: Sequence[np.ndarray], videos: Sequence[np.ndarray],
                    landmarks: Sequence[np.ndarray], valids: Optional[Sequence[np.ndarray]] = None,
                    audio_max_length: Optional[int] = N_SAMPLES) -> PackedBatch:
    """Host-side packing of per-utterance arrays.  Audio is cut to ``audio_max_length``
    (pad_or_trim's trim half; the pad half happens on the GPU).  Video is NOT cut here: the
    reference runs ``extract_lip_frames`` over the whole clip (preprocess/video_process.py:392-475)
    and trims the FEATURES afterwards (whisper_flamingo_ft_ami.py:299-302), so the last kept
    frames still see their own 12-frame smoothing window and landmark interpolation still reaches
    detections behind the cut.  The trim is applied on the output side: ``keep_frames`` of
    ``forward_collated`` / ``AVFrontEnd.split_lip(..., n_audio_samples=...)``."""
    assert len(audios) == len(videos) == len(landmarks)
    a_list, v_list, l_list, m_list = [], [], [], []
    for i in range(len(audios)):
        a = np.asarray(audios[i], dtype=np.float32).reshape(-1)
        if audio_max_length is not None:
            a = a[:audio_max_length]
        v = np.asarray(videos[i], dtype=np.uint8)
        lm = np.asarray(landmarks[i], dtype=np.float64)
        ok = np.ones(len(v), dtype=np.uint8) if valids is None else np.asarray(valids[i], dtype=np.uint8)
        a_list.append(a); v_list.append(v); l_list.append(lm); m_list.append(ok)
    a_off = np.concatenate([[0], np.cumsum([len(a) for a in a_list])]).astype(np.int64)
    c_off = np.concatenate([[0], np.cumsum([len(v) for v in v_list])]).astype(np.int64)
    return PackedBatch(torch.from_numpy(np.concatenate(a_list)), torch.from_numpy(a_off),
                       torch.from_numpy(np.concatenate(v_list)), torch.from_numpy(c_off),
                       torch.from_numpy(np.concatenate(l_list)), torch.from_numpy(np.concatenate(m_list)))


class AVFrontEnd:
    """Log-mel + lip-ROI features for a packed batch of utterances on one GPU."""

    def __init__(self, n_mels: int = 80, audio_max_length: int = N_SAMPLES, device=None,
                 want_gray: bool = True, want_lip_u8: bool = False, fused: bool = True,
                 image_crop_size: int = IMAGE_CROP_SIZE, image_mean: float = IMAGE_MEAN,
                 image_std: float = IMAGE_STD):
        _lib.require_cuda()
        _lib.load()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.n_mels = n_mels
        self.audio_max_length = audio_max_length
        self.want_gray = want_gray
        self.want_lip_u8 = want_lip_u8
        self.fused = fused
        self.crop = image_crop_size
        self.mean, self.std = image_mean, image_std
        self.filters = mel_filters(self.device, n_mels)
        self._bufs: Dict[str, torch.Tensor] = {}
        self._reuse = False

    @staticmethod
    def model_n_mels(model_name: str) -> int:
        """avsl/whisper_flamingo_ft_ami.py:212."""
        return 80 if "large-v3" not in model_name else 128

    def _buf(self, name: str, shape, dtype) -> torch.Tensor:
        """Output / scratch tensor: fresh unless the current call asked for ``reuse=True``."""
        if not self._reuse:
            return torch.empty(tuple(shape), dtype=dtype, device=self.device)
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    # ---------------------------------------------------------------- device-resident path
    def forward_device(self, batch: PackedBatch, padded_audio: Optional[torch.Tensor] = None,
                       mark=None, reuse: bool = False, lip_out: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """All inputs already on ``self.device``.  ``padded_audio`` [U, audio_max_length] makes the
        log-mel read an already padded matrix instead of the ragged ``batch.audio``.  ``mark(name)``
        (optional) is called after each stage is enqueued (bench.py records CUDA events there).
        Returns device tensors: mel [U,n_mels,F], lip [N,88,88,1], gray [N,H,W] (optional),
        lip_u8 [N,96,96] (optional).

        The returned tensors are freshly allocated.  ``reuse=True`` is the steady-state option: the
        outputs live in buffers owned by this object and the NEXT ``reuse=True`` call with the
        same shapes overwrites them in place (what ``capture`` and ``HostPipeline`` rely on) -- a
        loader that still holds batch i while batch i+1 is built must not use it.
        ``lip_out`` (float32 [N,crop,crop], a CUDA tensor or a PINNED host tensor) receives the lip
        features instead of a device buffer; with a pinned tensor the blend warps store across
        PCIe themselves, which overlaps with the footprint reads of the zero-copy mode."""
        U, L = batch.n_utts, self.audio_max_length
        mark = mark or (lambda name: None)
        self._reuse = bool(reuse)
        with torch.cuda.device(self.device):
            mark("start")
            mel = self._buf("mel", (U, self.n_mels, L // HOP_LENGTH), torch.float32)
            if padded_audio is None:
                # pad_or_trim is fused into the log-mel: the padded [U, L] matrix is never built
                log_mel_spectrogram_ragged(batch.audio, batch.audio_offsets, L, self.n_mels,
                                           filters=self.filters, out=mel)
            else:
                log_mel_spectrogram(padded_audio, self.n_mels, filters=self.filters, out=mel)
            mark("logmel")
            N, H, W = (int(s) for s in batch.frames.shape[:3])
            if lip_out is not None and not (lip_out.dtype == torch.float32 and lip_out.is_contiguous()
                                            and tuple(lip_out.shape) == (N, self.crop, self.crop)
                                            and (lip_out.is_cuda or lip_out.is_pinned())):
                raise ValueError(f"lip_out must be a contiguous float32 CUDA or pinned tensor of shape {(N, self.crop, self.crop)}")
            src = batch.frames
            gray = None
            bgr = batch.frames.dim() == 4
            if self.want_gray and bgr:
                gray = self._buf("gray", (N, H, W), torch.uint8)       # gray frames are a deliverable
                if not self.fused:
                    # two passes: stream-convert everything, then warp from the gray frames
                    _lib.call("avfe_bgr2gray_u8", _lib.ptr(batch.frames), N, H, W, _lib.ptr(gray), _lib.stream_ptr())
                    src = gray
                    mark("gray")
            res = LipBatch(gray if (self.fused and bgr) else None,
                             self._buf("lip_u8", (N, 96, 96), torch.uint8) if self.want_lip_u8 else None,
                             lip_out if lip_out is not None else self._buf("lip", (N, self.crop, self.crop), torch.float32),
                             None, None, batch.clip_offsets)
            # fused: one launch does the transform fits, the gray frames and the ROI warp (every
            # frame byte is read once)
            lip_roi_batch(src, batch.clip_offsets, batch.landmarks, batch.lm_valid,
                          want_gray=self.fused and gray is not None, want_u8=self.want_lip_u8,
                          crop=self.crop, image_mean=self.mean, image_std=self.std, out=res,
                          workspace=self._buf("lip_ws", (lip_workspace_bytes(N),), torch.uint8))
            mark("lip")
        out = {"mel": mel, "lip": res.lip_f32.unsqueeze(-1)}
        if gray is not None:
            out["gray"] = gray
        if res.lip_u8 is not None:
            out["lip_u8"] = res.lip_u8
        self._reuse = False
        return out

    # ---------------------------------------------------------------- the training batch, collated
    def forward_collated(self, batch: PackedBatch, clip_frames: Sequence[int], audio_lengths: Sequence[int],
                         spec_augment_config: Optional[str] = None, train: bool = False,
                         rng=None) -> Dict[str, torch.Tensor]:
        """The batch dict the reference's training step reads (``batch["input_ids"]``,
        ``batch["video"]``, ``batch["padding_mask"]``, avsl/whisper_flamingo_ft_ami.py:499-504,527),
        built on the GPU from a packed batch: what ``AmiVideoHFDataset.__getitem__`` (:187-313) does
        per sample -- pad_or_trim + log-mel, SpecAugment when ``train`` and a policy is set (:216-224),
        lip features trimmed to ``round(len(audio) / 16000 * 25)`` frames (:299-302; ``audio`` is the
        padded clip there) -- followed by the collator's zero padding and mask (:686).
        ``clip_frames`` / ``audio_lengths`` are the host-side frame and sample counts of the clips
        (``audio_frames_before_pad`` = samples // 160 bounds the time masks, :206)."""
        from .audio import spec_augment
        from .lips import lip_roi_collate
        U, L = batch.n_utts, self.audio_max_length
        self._reuse = False                       # every tensor of the returned dict is fresh
        keep_n = video_frames_for_audio(L)
        kept = [min(int(t), keep_n) for t in clip_frames]
        T_pad = max(max(kept), 1)
        with torch.cuda.device(self.device):
            mel = self._buf("mel", (U, self.n_mels, L // HOP_LENGTH), torch.float32)
            log_mel_spectrogram_ragged(batch.audio, batch.audio_offsets, L, self.n_mels, filters=self.filters, out=mel)
            if train and spec_augment_config:
                frames_before_pad = [min(int(n), L) // HOP_LENGTH for n in audio_lengths]
                spec_augment(mel, audio_frames=frames_before_pad, policy=spec_augment_config, rng=rng)
            keep = torch.full((U,), keep_n, dtype=torch.int64, device=self.device)
            col = lip_roi_collate(batch.frames, batch.clip_offsets, batch.landmarks, batch.lm_valid, T_pad=T_pad,
                                  keep_frames=keep, want_gray=self.want_gray, crop=self.crop,
                                  image_mean=self.mean, image_std=self.std)
        out = {"input_ids": mel, "video": col["video"], "padding_mask": col["padding_mask"]}
        if "gray" in col:
            out["gray"] = col["gray"]
        return out

    # ---------------------------------------------------------------- CUDA graph of one step
    def capture(self, batch: PackedBatch):
        """Capture the device-resident step on ``batch``'s buffers into a CUDA graph (the step is a
        handful of launches plus two memsets; replaying the graph removes their launch gaps, which
        matters for small batches).  Returns ``(graph, outputs)``: refill the tensors of ``batch``
        in place (same shapes), call ``graph.replay()``, read ``outputs`` (same tensors every time)."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.forward_device(batch, reuse=True)      # warm-up: buffers, filter packs, attributes
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.forward_device(batch, reuse=True)
        return graph, out

    # ---------------------------------------------------------------- host-buffer path (e2e)
    def zero_copy(self, batch: PackedBatch) -> bool:
        """Whether the host-buffer paths leave the frames in pinned host memory and let the lip
        kernel pull only each frame's ROI footprint across PCIe (about a seventh of the frame
        bytes): possible when no gray frames are wanted, which need every pixel on the device."""
        return (not self.want_gray) and (not batch.frames.is_cuda) and batch.frames.is_pinned()

    def forward_host(self, batch: PackedBatch, host_out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """``batch`` lives in (pinned) host memory: H2D copy of every input, the kernels, and a
        D2H copy of every deliverable: the features the reference's ``__getitem__`` returns (mel and
        lip) and, when ``want_gray``, the gray frames ``load_video`` returns.
        With ``want_gray=False`` the frames are not copied: see :meth:`zero_copy` (the batch must
        stay untouched until this call returns, as with any asynchronous copy)."""
        host_out = dict(host_out or {})
        res = self._run_host(batch, host_out)
        torch.cuda.current_stream(self.device).synchronize()
        return {k: host_out[k] for k in self.host_keys(res)}

    def _run_host(self, batch: PackedBatch, host_out: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Enqueue one host-buffer step on the current stream; ``host_out`` (pinned tensors, reused
        when their shapes fit, created otherwise) receives the results.  No synchronisation."""
        def pinned(name, shape, dtype):
            t = host_out.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
                t = host_out[name] = torch.empty(tuple(shape), dtype=dtype).pin_memory()
            return t
        zc = self.zero_copy(batch)
        dev = batch.to(self.device, non_blocking=True, frames_stay_on_host=zc)
        lip_out = None
        if zc and getattr(self, "lip_direct", True):      # the blend warps store the lip features across PCIe themselves: no device buffer, no D2H copy
            lip_out = pinned("lip", (int(batch.frames.shape[0]), self.crop, self.crop, 1), torch.float32).squeeze(-1)
        res = self.forward_device(dev, reuse=True, lip_out=lip_out)   # device results are copied out before the buffers are reused
        for k in self.host_keys(res):
            if res[k].is_cuda:
                pinned(k, res[k].shape, res[k].dtype).copy_(res[k], non_blocking=True)
        return res

    @staticmethod
    def host_keys(res: Dict[str, torch.Tensor]) -> List[str]:
        """The results the host-buffer paths copy back, in a fixed order."""
        return [k for k in ("mel", "lip", "gray", "lip_u8") if k in res]

    @staticmethod
    def split_lip(lip: torch.Tensor, clip_offsets, n_audio_samples: Optional[int] = None) -> List[torch.Tensor]:
        """Per-utterance [T_i,88,88,1] views of the packed lip tensor.  ``n_audio_samples`` (the
        padded audio length, ``audio_max_length``) applies the reference's trim to
        ``round(n_audio_samples / 16000 * 25)`` frames (whisper_flamingo_ft_ami.py:299-302)."""
        off = [int(x) for x in clip_offsets.tolist()]
        keep = None if n_audio_samples is None else video_frames_for_audio(n_audio_samples)
        return [lip[off[i]:(off[i + 1] if keep is None else min(off[i + 1], off[i] + keep))]
                for i in range(len(off) - 1)]


class HostPipeline:
    """Steady-state host-buffer path: ``depth`` front-ends, each with its own stream and device
    buffers, so that the H2D copy of batch i+1, the kernels of batch i and the D2H copy of batch
    i-1 overlap (PCIe is full duplex and the copy engines run beside the SMs).  Every batch still
    pays its full H2D of inputs and D2H of mel + lip (+ gray when ``want_gray``); only the waiting
    is overlapped.  With ``want_gray=False`` the frames stay in pinned host memory and the lip
    kernel reads just the ROI footprints through the mapped pointer (``AVFrontEnd.zero_copy``);
    a submitted batch must not be modified before ``result(i)`` returns."""

    def __init__(self, depth: int = 2, **frontend_kwargs):
        self.fes = [AVFrontEnd(**frontend_kwargs) for _ in range(depth)]
        dev = self.fes[0].device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        self.events = [None] * depth
        self.outs: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self.device = dev

    def submit(self, i: int, batch: PackedBatch) -> None:
        """Enqueue batch ``i`` (pinned host tensors) on slot ``i % depth``; returns immediately."""
        k = i % len(self.fes)
        if self.events[k] is not None:
            self.events[k].synchronize()              # slot's previous batch fully drained
        fe, st = self.fes[k], self.streams[k]
        st.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(st):
            if self.outs[k] is None:
                self.outs[k] = {}
            fe._run_host(batch, self.outs[k])            # the slot's event guards its buffers
            ev = torch.cuda.Event()
            ev.record(st)
            self.events[k] = ev

    def result(self, i: int) -> Dict[str, torch.Tensor]:
        """Host features of batch ``i`` (blocks until its D2H copy has landed); valid until the
        slot is submitted to again."""
        k = i % len(self.fes)
        self.events[k].synchronize()
        return self.outs[k]

    def drain(self) -> None:
        for ev in self.events:
            if ev is not None:
                ev.synchronize()


def bind_to_gpu_numa_node(device_index: int) -> Optional[List[int]]:
    """Pin the calling process to the CPUs closest to GPU ``device_index`` (NVML's ideal affinity),
    so that the pinned staging buffers it allocates afterwards are first-touched on the NUMA node
    the GPU hangs off and host-to-device copies do not cross the socket interconnect.  One process
    per GPU calls this once, before allocating pinned memory.  Returns the CPU list, or ``None`` when
    NVML or the scheduler call is unavailable (nothing is changed then)."""
    try:
        import os

        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def shard(n_items: int, rank: int, world_size: int) -> np.ndarray:
    """Indices of the utterances rank ``rank`` of ``world_size`` processes owns."""
    if not (0 <= rank < world_size):
        raise ValueError("rank must be in [0, world_size)")
    return np.arange(rank, n_items, world_size)


def shard_balanced(durations, rank: int, world_size: int) -> np.ndarray:
    """Length-balanced variant of ``shard`` (the cost of an utterance is proportional to its
    duration: SURVEY 8(e)): longest-processing-time-first assignment of the utterances to the
    ranks, computed identically by every rank from the duration list alone -- no communication.
    Returns the sorted indices rank ``rank`` owns; the ranks' sets partition ``range(len(durations))``
    and their total durations differ by at most the longest utterance."""
    if not (0 <= rank < world_size):
        raise ValueError("rank must be in [0, world_size)")
    d = np.asarray(durations, dtype=np.float64).reshape(-1)
    order = np.argsort(-d, kind="stable")                   # longest first, ties by index: deterministic
    load = np.zeros(world_size, dtype=np.float64)
    owner = np.empty(len(d), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))                            # first of the least-loaded ranks
        owner[i] = r
        load[r] += d[i]
    return np.flatnonzero(owner == rank)


def aggregate_rank_stats(elapsed_ms: float, audio_seconds: float, alg_bytes: float, launches: float = 0.0,
                         device="cpu"):
    """Job-level numbers from per-rank ones: time = MAX over ranks, work = SUM over ranks.
    This reporting reduction is the only collective in the package and is off the hot path.
    Returns (elapsed_ms_max, audio_seconds_total, alg_bytes_total, launches_total)."""
    import torch.distributed as dist
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    w = torch.tensor([audio_seconds, alg_bytes, launches], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    a, b, c = (float(v) for v in w.tolist())
    return float(t.item()), a, b, c


def algorithmic_bytes(n_utts: int, n_frames: int, H: int = 224, W: int = 224, n_mels: int = 80,
                      audio_len: int = N_SAMPLES, crop: int = IMAGE_CROP_SIZE) -> int:
    """per utterance 4*L + 4*n_mels*(L/160); per frame H*W*3 + 68*2*8 + H*W +
    crop*crop*4."""
    per_utt = 4 * audio_len + 4 * n_mels * (audio_len // HOP_LENGTH)
    per_frame = H * W * 3 + 68 * 2 * 8 + H * W + crop * crop * 4
    return n_utts * per_utt + n_frames * per_frame
