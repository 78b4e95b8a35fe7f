"""Generate the committed golden vectors (run in the build container, where /root/reference and
transformers / cv2 are available):

    python tests/golden/make_golden.py

* logmel_golden.npz    transformers.WhisperFeatureExtractor — the reference's literal call at
                       avsl/whisper_ft.py:347-350 — on seeded clips (30 s padded path, selected
                       frames; short unpadded clips, full output), n_mels 80 and 128, plus its
                       mel filterbanks.
* gray_golden.npz      cv2.cvtColor(BGR2GRAY) (preprocess/video_process.py:214) on a seeded image
                       and on a sweep that hits every rounding boundary class.
* video_feats_golden.npz  the reference's own load_video_feats_from_decord_reader
                       (utils/hf_video_utils.py:73-145), imported by file path from
                       /root/reference and fed a duck-typed reader.
The similarity-fit / warp path has no golden: scikit-image is not installable here
("parity unpinned", see oracle/lips.py).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from avsl_b200 import synth  # noqa: E402  (seeded generators only; no GPU needed)

FRAME_SEL = np.r_[0:64, 1000:1032, 2936:3000]


def logmel():
    from transformers import WhisperFeatureExtractor
    out = {"frame_sel": FRAME_SEL}
    clips = {
        "noise30": synth.audio_clip(480000, 3407),
        "chirp30": synth.chirp_silence_clip(480000),
        "short7s": synth.audio_clip(112000, 11) * 3.0,      # padded to 30 s by the extractor
    }
    shorts = {
        "s1": synth.audio_clip(16000, 5),
        "s2": synth.chirp_silence_clip(24000),
        "s3": synth.audio_clip(4321, 7),                     # ragged length, not a hop multiple
    }
    for n_mels in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=n_mels)
        out[f"filters_{n_mels}"] = np.ascontiguousarray(fe.mel_filters.T.astype(np.float32))
        for name, a in clips.items():
            ref = fe(a, sampling_rate=16000, return_tensors="np").input_features[0]
            assert ref.shape == (n_mels, 3000)
            out[f"{name}_{n_mels}"] = ref[:, FRAME_SEL].astype(np.float32)
        for name, a in shorts.items():
            ref = fe(a, sampling_rate=16000, return_tensors="np", padding=False,
                     truncation=False).input_features[0]
            out[f"{name}_{n_mels}"] = ref.astype(np.float32)
    for name, a in {**clips, **shorts}.items():
        out[f"audio_{name}"] = a if len(a) < 30000 else np.zeros(0, np.float32)  # long ones are re-generated from the seed
    np.savez_compressed(os.path.join(HERE, "logmel_golden.npz"), **out)
    print("logmel:", {k: v.shape for k, v in out.items()})


def gray():
    import cv2
    rng = np.random.default_rng(3407)
    img = rng.integers(0, 256, size=(48, 64, 3), dtype=np.uint8)
    # sweep: all (b, g, r) with two channels on a coarse lattice and one dense
    lat = np.array([0, 1, 2, 63, 64, 127, 128, 129, 200, 254, 255], dtype=np.uint8)
    dense = np.arange(256, dtype=np.uint8)
    sweep = np.stack(np.meshgrid(lat, dense, lat, indexing="ij"), axis=-1).reshape(1, -1, 3)
    np.savez_compressed(os.path.join(HERE, "gray_golden.npz"), img=img,
                        img_gray=cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), sweep=sweep,
                        sweep_gray=cv2.cvtColor(np.ascontiguousarray(sweep), cv2.COLOR_BGR2GRAY))
    print("gray: ok", sweep.shape)


def video_feats():
    path = "/root/reference/utils/hf_video_utils.py"
    spec = importlib.util.spec_from_file_location("ref_hf_video_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class FakeVideoReader:
        def __init__(self, frames):
            self.frames = frames

        def __len__(self):
            return len(self.frames)

        def get_batch(self, idx):
            return self.frames[idx]

    rng = np.random.default_rng(3407)
    roi = rng.integers(0, 256, size=(3, 96, 96), dtype=np.uint8)
    roi[0, :8, :8] = 0
    roi[0, 8:16, :8] = 255
    feats = mod.load_video_feats_from_decord_reader(FakeVideoReader(roi[..., None]), train=False,
                                                    image_crop_size=88, image_mean=0.421, image_std=0.165)
    feats = feats.astype(np.float32)      # avsl/whisper_flamingo_ft_ami.py:286
    assert feats.shape == (3, 88, 88, 1)
    # all 256 levels through the same function (one 96x96 frame holding every value)
    lv = np.resize(np.arange(256, dtype=np.uint8), (1, 96, 96))
    lv_feats = mod.load_video_feats_from_decord_reader(FakeVideoReader(lv[..., None]), image_crop_size=88,
                                                       image_mean=0.421, image_std=0.165).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "video_feats_golden.npz"), roi=roi, feats=feats, levels=lv,
                        levels_feats=lv_feats)
    print("video_feats:", feats.shape, feats.dtype, float(feats.min()), float(feats.max()))


if __name__ == "__main__":
    logmel()
    gray()
    video_feats()
