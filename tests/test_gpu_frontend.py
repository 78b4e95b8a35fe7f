"""GPU: the batched AV front-end (host-buffer path and device path) vs per-utterance oracle."""
import numpy as np
import pytest
import torch

import avsl_b200 as A
from avsl_b200 import synth
from oracle import lips as OL
from oracle import logmel as OM

pytestmark = pytest.mark.gpu


def _utts(n, seed=0):
    rng = np.random.default_rng(seed)
    durs = [0.36, 1.0, 2.52][:n] if n <= 3 else list(rng.uniform(0.3, 2.0, n).round(2))
    audios, vids, lms, vals = [], [], [], []
    for i, d in enumerate(durs):
        T = max(1, int(round(d * 25)))
        audios.append(synth.audio_clip(int(d * 16000), seed + i))
        f, lm, v = synth.video_clip(T, 96, 128, seed=seed + 10 + i, invalid_frac=0.1)
        vids.append(f); lms.append(lm); vals.append(v)
    return audios, vids, lms, vals


def test_forward_host_matches_per_utterance_oracle():
    audios, vids, lms, vals = _utts(3)
    L = 48000
    fe = A.AVFrontEnd(n_mels=80, audio_max_length=L, want_lip_u8=True)
    batch = A.pack_utterances(audios, vids, lms, vals, audio_max_length=L).pin()
    out = fe.forward_host(batch)
    assert out["mel"].shape == (3, 80, 300)
    lips = fe.split_lip(out["lip"], batch.clip_offsets)
    mf = A.mean_face_landmarks()
    for i in range(3):
        ref_mel = OM.log_mel_spectrogram(OM.pad_or_trim(audios[i], L), 80)
        assert (out["mel"][i] - ref_mel).abs().max().item() <= 1e-4
        lst = [lms[i][k] if vals[i][k] else None for k in range(len(lms[i]))]
        roi, _, _ = OL.extract_lip_frames_from_arrays(OL.bgr2gray(vids[i]), lst, mf)
        ref_feats = OL.video_feats_from_u8(roi)
        assert lips[i].shape == ref_feats.shape and lips[i].dtype == torch.float32
        assert np.abs(lips[i].numpy() - ref_feats).max() <= (1 / 255) / 0.165 + 1e-6


def test_device_path_equals_host_path_and_is_repeatable():
    audios, vids, lms, vals = _utts(5, seed=3)
    fe = A.AVFrontEnd(n_mels=128, audio_max_length=32000)
    batch = A.pack_utterances(audios, vids, lms, vals, audio_max_length=32000)
    host = fe.forward_host(batch.pin())
    dev = fe.forward_device(batch.to("cuda"))
    assert torch.equal(dev["mel"].cpu(), host["mel"]) and torch.equal(dev["lip"].cpu(), host["lip"])
    np.testing.assert_array_equal(dev["gray"].cpu().numpy(), OL.bgr2gray(np.concatenate(vids)))
    assert A.AVFrontEnd.model_n_mels("large-v3") == 128 and A.AVFrontEnd.model_n_mels("large-v2") == 80


def test_launch_counter_counts_kernels():
    from avsl_b200 import _lib
    before = _lib.launch_count()
    A.bgr2gray(np.zeros((2, 32, 32, 3), np.uint8))
    assert _lib.launch_count() > before


def test_host_pipeline_matches_serial_path():
    """Two slots in flight give the same bytes as the step-at-a-time path, for every step."""
    batches = []
    for seed in (0, 3, 7):
        audios, vids, lms, vals = _utts(4, seed=seed)
        batches.append(A.pack_utterances(audios, vids, lms, vals, audio_max_length=32000).pin())
    fe = A.AVFrontEnd(n_mels=80, audio_max_length=32000)
    ref = [{k: v.clone() for k, v in fe.forward_host(b).items()} for b in batches]
    pipe = A.HostPipeline(depth=2, n_mels=80, audio_max_length=32000)
    for rounds in range(2):
        for i, b in enumerate(batches):
            pipe.submit(i, b)
            if i >= 1:                                     # consume with one step of lag
                got = pipe.result(i - 1)
                assert torch.equal(got["mel"], ref[i - 1]["mel"]) and torch.equal(got["lip"], ref[i - 1]["lip"])
        got = pipe.result(len(batches) - 1)
        assert torch.equal(got["mel"], ref[-1]["mel"]) and torch.equal(got["lip"], ref[-1]["lip"])
    pipe.drain()


def test_cuda_graph_replay_matches_eager():
    """One step captured into a CUDA graph: replay on refilled input buffers gives the eager result."""
    def _batch(seed):
        audios, vids, lms, vals = _utts(3, seed)          # same durations for every seed: same shapes
        return A.pack_utterances(audios, vids, lms, vals, audio_max_length=32000)
    b1 = _batch(seed=1).to("cuda")
    b2 = _batch(seed=2).to("cuda")
    fe = A.AVFrontEnd(n_mels=80, audio_max_length=32000, want_gray=True)
    eager2 = {k: v.clone() for k, v in fe.forward_device(b2).items()}
    graph, out = fe.capture(b1)
    graph.replay()
    torch.cuda.synchronize()
    eager1 = {k: v.clone() for k, v in A.AVFrontEnd(n_mels=80, audio_max_length=32000, want_gray=True).forward_device(b1).items()}
    for k in eager1:
        assert torch.equal(out[k], eager1[k]), k
    for name in A.PackedBatch.FIELDS:                     # same shapes: refill in place and replay
        getattr(b1, name).copy_(getattr(b2, name))
    graph.replay()
    torch.cuda.synchronize()
    for k in eager2:
        assert torch.equal(out[k], eager2[k]), k


def test_forward_collated_is_getitem_plus_collator():
    """forward_collated == per-sample pad_or_trim + log-mel (+ SpecAugment with the same bands) and
    lip features trimmed to round(L/16000*25) frames, then the collator's padding and mask."""
    from oracle import collate as OC
    from oracle import specaug as OS
    audios, vids, lms, vals = _utts(3)
    L = 16000                                            # 1 s: keeps 25 frames, the 2.52 s clip (63 frames) is trimmed
    fe = A.AVFrontEnd(n_mels=80, audio_max_length=L)
    batch = A.pack_utterances(audios, vids, lms, vals, audio_max_length=None)
    frames = [len(v) for v in vids]
    lens = [len(a) for a in audios]
    out = fe.forward_collated(batch.to("cuda"), frames, lens, "ls-double", train=True, rng=np.random.default_rng(9))
    out = {k: v.clone() for k, v in out.items()}         # the front-end reuses its output buffers
    plain = fe.forward_device(batch.to("cuda"))
    lips = fe.split_lip(plain["lip"].cpu(), batch.clip_offsets)
    keep = OC.video_frames_for_audio(L)
    assert keep == 25
    ref = OC.collate_video([OC.trim_video(x.numpy(), L) for x in lips])
    assert out["video"].shape == (3, 1, 25, 88, 88)
    np.testing.assert_array_equal(out["video"].cpu().numpy(), ref["video"])
    np.testing.assert_array_equal(out["padding_mask"].cpu().numpy(), ref["padding_mask"])
    bands = A.spec_augment_bands([min(n, L) // 160 for n in lens], 80, "ls-double", np.random.default_rng(9))
    np.testing.assert_array_equal(out["input_ids"].cpu().numpy(), OS.apply_bands(plain["mel"].cpu().numpy(), bands))
    # eval: no augmentation
    ev = fe.forward_collated(batch.to("cuda"), frames, lens, "ls-double", train=False)["input_ids"].clone()
    assert torch.equal(ev, fe.forward_device(batch.to("cuda"))["mel"])


def test_trim_happens_after_the_lip_path_not_before():
    """A clip longer than the kept frame count: the reference runs extract_lip_frames on
    the WHOLE clip and trims the features afterwards (whisper_flamingo_ft_ami.py:299-302), so the last
    kept frames keep their own 12-frame window and a failed detection near the cut is interpolated
    towards the detection behind it.  pack_utterances with an audio_max_length must give that."""
    L = 16000                                            # keeps 25 frames
    T = 60
    frames, lm, valid = synth.video_clip(T, 96, 128, seed=77, invalid_frac=0.0)
    valid[:] = 1
    valid[22:29] = 0                                     # a gap straddling the cut: filled from frame 29, behind it
    audio = synth.audio_clip(48000, 5)
    batch = A.pack_utterances([audio], [frames], [lm], [valid], audio_max_length=L)
    assert int(batch.clip_offsets[-1]) == T              # video is not pre-trimmed
    assert int(batch.audio_offsets[-1]) == L
    fe = A.AVFrontEnd(n_mels=80, audio_max_length=L, want_lip_u8=True)
    dev = batch.to("cuda")
    out = fe.forward_device(dev)
    lst = [lm[k] if valid[k] else None for k in range(T)]
    roi, _, _ = OL.extract_lip_frames_from_arrays(OL.bgr2gray(frames), lst, A.mean_face_landmarks())
    ref = OL.trim_video_to_audio(OL.video_feats_from_u8(roi), L)
    assert ref.shape[0] == 25
    got = fe.split_lip(out["lip"].cpu(), batch.clip_offsets, n_audio_samples=L)[0].numpy()
    assert got.shape == ref.shape
    lv = np.rint(np.abs(got - ref) * 0.165 * 255)
    assert lv.max() <= 1 and (lv != 0).mean() < 1e-3
    col = fe.forward_collated(dev, [T], [len(audio)])
    assert col["video"].shape == (1, 1, 25, 88, 88)
    np.testing.assert_array_equal(col["video"][0, 0].cpu().numpy(), got[..., 0])
    # cutting the video first (what round 1's pack_utterances did) is a different computation
    pre = A.pack_utterances([audio], [frames[:25]], [lm[:25]], [valid[:25]], audio_max_length=L).to("cuda")
    cut = fe.forward_device(pre)["lip"].cpu().numpy()
    assert not np.array_equal(cut, got)


def test_outputs_are_fresh_unless_reuse_is_requested():
    """forward_device / forward_collated must not hand out buffers the next call overwrites."""
    a1, v1, l1, m1 = _utts(3, seed=1)
    a2, v2, l2, m2 = _utts(3, seed=2)
    b1 = A.pack_utterances(a1, v1, l1, m1, audio_max_length=32000).to("cuda")
    b2 = A.pack_utterances(a2, v2, l2, m2, audio_max_length=32000).to("cuda")
    fe = A.AVFrontEnd(n_mels=80, audio_max_length=32000)
    o1 = fe.forward_device(b1)
    keep = {k: v.clone() for k, v in o1.items()}
    o2 = fe.forward_device(b2)
    for k in o1:
        assert o1[k].data_ptr() != o2[k].data_ptr(), k
        assert torch.equal(o1[k], keep[k]), k
    c1 = fe.forward_collated(b1, [len(v) for v in v1], [len(a) for a in a1])
    keepc = {k: v.clone() for k, v in c1.items()}
    fe.forward_collated(b2, [len(v) for v in v2], [len(a) for a in a2])
    for k in c1:
        assert torch.equal(c1[k], keepc[k]), k
    # opt-in reuse: same storage, overwritten in place
    r1 = fe.forward_device(b1, reuse=True)
    p = {k: v.data_ptr() for k, v in r1.items()}
    r2 = fe.forward_device(b2, reuse=True)
    assert all(r2[k].data_ptr() == p[k] for k in p)
    assert torch.equal(r2["mel"], o2["mel"]) and torch.equal(r2["lip"], o2["lip"])


def test_zero_copy_frames_give_the_same_features():
    """want_gray=False host path: the frames stay in pinned host memory and the lip kernel reads the
    ROI footprints through the mapped pointer -- same bytes as the device-resident path, for
    interior ROIs, ROIs hanging over the frame border (global taps) and a clip without detections."""
    audios, vids, lms, vals = _utts(4, seed=5)
    lms[1] = lms[1] + np.array([40.0, 30.0])            # mouth near the frame edge: footprint not interior
    vals[2][:] = 0                                      # no detection at all: zero ROI
    batch = A.pack_utterances(audios, vids, lms, vals, audio_max_length=32000).pin()
    ref = A.AVFrontEnd(n_mels=80, audio_max_length=32000, want_gray=True).forward_host(batch)
    fe = A.AVFrontEnd(n_mels=80, audio_max_length=32000, want_gray=False)
    assert fe.zero_copy(batch)
    out = fe.forward_host(batch)
    assert sorted(out) == ["lip", "mel"]
    assert torch.equal(out["lip"], ref["lip"]) and torch.equal(out["mel"], ref["mel"])
    pipe = A.HostPipeline(depth=2, n_mels=80, audio_max_length=32000, want_gray=False)
    for i in range(3):
        pipe.submit(i, batch)
    assert torch.equal(pipe.result(2)["lip"], ref["lip"])
    pipe.drain()
    # the C ABI contract: host frames only without gray output
    dev = batch.to("cuda", frames_stay_on_host=True)
    with pytest.raises(ValueError, match="want_gray"):
        A.lip_roi_batch(dev.frames, dev.clip_offsets, dev.landmarks, dev.lm_valid, want_gray=True)
    with pytest.raises(ValueError, match="PINNED"):
        A.lip_roi_batch(torch.zeros((2, 8, 8, 3), dtype=torch.uint8), dev.clip_offsets, dev.landmarks, dev.lm_valid,
                        want_gray=False)
