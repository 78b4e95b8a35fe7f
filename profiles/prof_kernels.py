"""Small driver for ncu captures: runs one hot-path kernel group a few times on BASELINE-sized
synthetic input.  Usage (on the GPU box):

    python profiles/prof_kernels.py logmel|lip|fuse|all [--iters 3]

Numbers printed by a run under ncu are never bench values; bench.py is the only source of those.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import avsl_b200 as A  # noqa: E402
from avsl_b200 import lips as L  # noqa: E402
from avsl_b200 import synth  # noqa: E402


def run_logmel(iters, n_mels=80, batch=64):
    a = synth.audio_batch(batch, 480000, 3407, device="cuda")
    out = torch.empty((batch, n_mels, 3000), device="cuda")
    for _ in range(iters):
        A.log_mel_spectrogram(a, n_mels, out=out)
    torch.cuda.synchronize()


def run_lip(iters, clips=32, T=250):
    N = clips * T
    frames = synth.video_frames_cuda(N, 224, 224, device="cuda")
    lms, vals = zip(*[synth.landmarks_for_clip(T, seed=3407 + i) for i in range(clips)])
    lm = torch.from_numpy(np.concatenate(lms)).cuda()
    valid = torch.from_numpy(np.concatenate(vals)).cuda()
    off = torch.arange(0, N + 1, T, dtype=torch.int64, device="cuda")
    res = None
    for _ in range(iters):
        res = L.lip_roi_batch(frames, off, lm, valid, want_gray=True, want_f32=True, out=res)
    torch.cuda.synchronize()


def run_fuse(iters):
    fa, fv, mask = synth.fusion_inputs(64, 1024, 750, device="cuda")
    for mode in ("concat", "add"):
        out = None
        for _ in range(iters):
            out = A.fuse_modalities(fa, fv, mask, mode, out=out)
    torch.cuda.synchronize()


def run_logfbank(iters, batch=64):
    a = synth.audio_batch(batch, 480000, 3407, device="cuda").reshape(-1)
    plan = A.LogfbankPlan(np.arange(batch + 1, dtype=np.int64) * 480000, 4, 26, a.device)
    for _ in range(iters):
        A.logfbank_batch(a, plan=plan, normalize=True)
    torch.cuda.synchronize()


def run_noise(iters, batch=64):
    wav = synth.audio_batch(batch, 480000, 3407, device="cuda").reshape(-1) * 8000.0
    nz = synth.audio_batch(batch, 160000, 3408, device="cuda").reshape(-1) * 3000.0
    plan = A.NoisePlan(np.arange(batch + 1, dtype=np.int64) * 480000, np.arange(batch + 1, dtype=np.int64) * 160000,
                       10, wav.device)
    for _ in range(iters):
        A.add_noise_batch(wav, noise=nz, plan=plan)
    torch.cuda.synchronize()


def run_fuse_ln(iters):
    fa, fv, mask = synth.fusion_inputs(64, 1024, 750, device="cuda")
    w, b = torch.ones(2048, device="cuda"), torch.zeros(2048, device="cuda")
    out = None
    for _ in range(iters):
        out = A.fuse_transpose_layernorm(fa, fv, mask, "concat", w, b, out=out)
    out = None
    fa, fv = fa.half(), fv.half()
    for _ in range(iters):
        out = A.fuse_transpose_layernorm(fa, fv, mask, "concat", w, b, out=out)
    torch.cuda.synchronize()


def run_fuse_ln_tma(iters):
    """the padded-pitch layout ([B, C, 752] allocation): tensor-map TMA tile fill"""
    for dtype in (torch.float32, torch.float16):
        fa0, fv0, mask = synth.fusion_inputs(64, 1024, 750, device="cuda")
        fa, fv = A.alloc_features(64, 1024, 750, dtype, "cuda"), A.alloc_features(64, 1024, 750, dtype, "cuda")
        fa.copy_(fa0.to(dtype)); fv.copy_(fv0.to(dtype))
        w, b = torch.ones(2048, device="cuda"), torch.zeros(2048, device="cuda")
        out = None
        for _ in range(iters):
            out = A.fuse_transpose_layernorm(fa, fv, mask, "concat", w, b, out=out)
    torch.cuda.synchronize()


def run_proj(iters):
    """post_extract_proj behind the folded LayerNorm: pep_stats_kernel + pep_gemm2_kernel (tcgen05)"""
    B, C, T, D = 64, 1024, 750, 1024
    g = torch.Generator(device="cuda").manual_seed(0)
    fa, fv = A.alloc_features(B, C, T, torch.float16, "cuda"), A.alloc_features(B, C, T, torch.float16, "cuda")
    fa.copy_(torch.randn(B, C, T, generator=g, device="cuda").half()); fv.copy_(torch.randn(B, C, T, generator=g, device="cuda").half())
    W = torch.randn(D, 2 * C, generator=g, device="cuda") / (2 * C) ** 0.5
    folded = A.FoldedProjection(W, torch.zeros(D, device="cuda"), torch.ones(2 * C, device="cuda"), torch.zeros(2 * C, device="cuda"), torch.float16)
    out = torch.empty((B, T, D), dtype=torch.float16, device="cuda")
    for _ in range(iters):
        A.fuse_layernorm_project(fa, fv, None, folded, out=out)
    torch.cuda.synchronize()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["logmel", "lip", "fuse", "fuse_ln", "fuse_ln_tma", "proj", "logfbank", "noise", "all"])
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    if a.which in ("logmel", "all"):
        run_logmel(a.iters)
    if a.which in ("lip", "all"):
        run_lip(a.iters)
    if a.which in ("fuse", "all"):
        run_fuse(a.iters)
    if a.which in ("logfbank", "all"):
        run_logfbank(a.iters)
    if a.which in ("noise", "all"):
        run_noise(a.iters)
    if a.which in ("fuse_ln", "all"):
        run_fuse_ln(a.iters)
    if a.which in ("fuse_ln_tma", "all"):
        run_fuse_ln_tma(a.iters)
    if a.which in ("proj", "all"):
        run_proj(a.iters)
    print("done", a.which)
