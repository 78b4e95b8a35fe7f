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
                       /root/reference and fed a duck-typed reader: the uint8 1-channel case, the RGB
                       frames decord really returns (bright and all-dark stacks) and frames smaller
                       than the crop (the cv2.resize fallback, with IPP on -- as shipped -- and off).
* noise_golden.npz      the reference's own add_noise (preprocess/audio_process.py:110-150), taken
                       out of its module by function name (the module's top-level imports --
                       librosa, python_speech_features -- are not installable here) and run on
                       seeded int16-scale and [-1, 1] waveforms: tiled / cut / equal-length noise,
                       both clipping branches, lengths below 8 and around 128, one 30 s clip
                       (slices + SHA-256 of the whole output).
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
    # what the call site really feeds it (avsl/whisper_flamingo_ft_ami.py:279-286): decord's RGB frames
    def run(frames, crop=88):
        return mod.load_video_feats_from_decord_reader(FakeVideoReader(frames), image_crop_size=crop,
                                                       image_mean=0.421, image_std=0.165).astype(np.float32)
    import cv2
    extra = {}
    rgb = rng.integers(0, 256, size=(2, 96, 96, 3), dtype=np.uint8)
    extra["rgb"], extra["rgb_feats"] = rgb, run(rgb)
    dark = rng.integers(0, 2, size=(2, 96, 96, 3), dtype=np.uint8)       # stack maximum <= 1.0: no /255, float64 kept
    dark[..., 1] = 0
    assert np.dot(dark[..., :3], [0.2989, 0.5870, 0.1140]).max() <= 1.0
    extra["dark"], extra["dark_feats"] = dark, run(dark)
    # frames smaller than the crop -> cv2.resize; stored for cv2 as shipped (IPP) and with IPP off
    for name, shape in (("small_rgb", (2, 64, 64, 3)), ("small_gray", (2, 64, 64, 1)), ("small_tall", (2, 120, 64, 3)),
                        ("small_wide", (1, 80, 100, 3))):
        fr = rng.integers(0, 256, size=shape, dtype=np.uint8)
        extra[name] = fr
        cv2.ipp.setUseIPP(True)
        extra[name + "_feats_ipp"] = run(fr)
        cv2.ipp.setUseIPP(False)
        extra[name + "_feats_noipp"] = run(fr)
        cv2.ipp.setUseIPP(True)
    extra["cv2_build_has_ipp"] = np.array("Intel IPP:" in cv2.getBuildInformation() and "Intel IPP:                   NO" not in cv2.getBuildInformation())
    np.savez_compressed(os.path.join(HERE, "video_feats_golden.npz"), roi=roi, feats=feats, levels=lv,
                        levels_feats=lv_feats, **extra)
    print("video_feats:", feats.shape, feats.dtype, float(feats.min()), float(feats.max()))


def noise_cases():
    """name -> (clean f32, noise f32, snr); int16-scale unless said otherwise.  Shared with the tests
    (tests/test_oracle_noise.py regenerates the 30 s case from its seed)."""
    rng = np.random.default_rng(3407)

    def wav(n, amp):
        return np.clip(np.rint(rng.standard_normal(n) * amp), -32768, 32767).astype(np.int16).astype(np.float32)

    cases = {
        "tile": (wav(24000, 3000), wav(7345, 800), 10),
        "cut": (wav(12000, 2500), wav(26000, 4000), 0),
        "equal": (wav(9000, 1000), wav(9000, 1000), -5),
        "multiple": (wav(3 * 4096, 2000), wav(4096, 500), 7.5),
        "tiny": (wav(5, 3000), wav(3, 1000), 5),
        "n130": (wav(130, 3000), wav(129, 1000), 3),
        "n8": (wav(8, 3000), wav(300, 1000), 20),
        "unit_range": ((rng.standard_normal(8000) * 0.1).astype(np.float32),
                       (rng.standard_normal(3000) * 0.1).astype(np.float32), 10),
    }
    hot = wav(12000, 9000)
    hot[2345] = 32000.0
    cases["clip_pos"] = (hot, wav(7000, 6000), -6)
    cold = wav(12000, 9000)
    cold[777] = -32700.0
    cold[778] = -32768.0
    cases["clip_neg"] = (cold, wav(13000, 6000), -6)
    return cases


def noise_long_case():
    rng = np.random.default_rng(99)
    clean = rng.integers(-12000, 12001, size=480000).astype(np.float32)
    noise = rng.integers(-3000, 3001, size=160007).astype(np.float32)
    return clean, noise, 4


NOISE_SEL = np.r_[0:256, 160000:160256, 479744:480000]


def noise_mix():
    import ast
    import hashlib
    path = "/root/reference/preprocess/audio_process.py"
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "add_noise"]
    ns = {"np": np}
    exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), ns)   # the reference's function, unmodified
    add_noise = ns["add_noise"]
    out = {}
    for name, (clean, noise, snr) in noise_cases().items():
        small = (lambda a: a.astype(np.int16) if name != "unit_range" else a)   # int16-valued: store as int16
        out[f"{name}_clean"], out[f"{name}_noise"], out[f"{name}_snr"] = small(clean), small(noise), np.float64(snr)
        out[f"{name}_mixed"] = add_noise(clean.copy(), noise.copy(), snr)
        assert out[f"{name}_mixed"].dtype == np.int16
    clean, noise, snr = noise_long_case()
    mixed = add_noise(clean, noise, snr)
    out["long_sel"] = NOISE_SEL
    out["long_mixed_sel"] = mixed[NOISE_SEL]
    out["long_sha256"] = np.frombuffer(hashlib.sha256(mixed.tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "noise_golden.npz"), **out)
    print("noise:", {k: (v.shape, int(np.abs(v).max()) if v.size else 0) for k, v in out.items() if k.endswith("_mixed")})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "noise":
        noise_mix()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "video_feats":
        video_feats()
        sys.exit(0)
    logmel()
    gray()
    video_feats()
    noise_mix()
