"""GPU parity of the SNR noise mixing (avfe_add_noise through avsl_b200.audio) -- int16 results,
so the bar is bit-exact: against the golden outputs of the reference's own add_noise
(tests/golden/noise_golden.npz) and against the oracle on seeded ragged batches."""
import hashlib
import os
import sys

import numpy as np
import pytest

from oracle import noise as ON

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as MG  # noqa: E402  (seeded input generators only)

pytestmark = pytest.mark.gpu
CASES = ["tile", "cut", "equal", "multiple", "tiny", "n130", "n8", "unit_range", "clip_pos", "clip_neg"]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "noise_golden.npz"))


@pytest.mark.parametrize("name", CASES)
def test_single_clip_equals_reference_output(golden, name):
    import avsl_b200 as A
    out = A.add_noise(golden[f"{name}_clean"], golden[f"{name}_noise"], float(golden[f"{name}_snr"]))
    assert out.dtype == np.int16
    np.testing.assert_array_equal(out, golden[f"{name}_mixed"])


def test_long_clip_equals_reference_output(golden):
    import avsl_b200 as A
    clean, noise, snr = MG.noise_long_case()
    out = A.add_noise(clean, noise, snr)
    np.testing.assert_array_equal(out[golden["long_sel"]], golden["long_mixed_sel"])
    assert hashlib.sha256(out.tobytes()).digest() == golden["long_sha256"].tobytes()


def test_packed_batch_of_all_golden_cases(golden):
    import torch
    import avsl_b200 as A
    cl = [golden[f"{n}_clean"].astype(np.float32) for n in CASES]
    nz = [golden[f"{n}_noise"].astype(np.float32) for n in CASES]
    co = np.cumsum([0] + [len(c) for c in cl])
    no = np.cumsum([0] + [len(c) for c in nz])
    snr = [float(golden[f"{n}_snr"]) for n in CASES]
    c, z = torch.from_numpy(np.concatenate(cl)).cuda(), torch.from_numpy(np.concatenate(nz)).cuda()
    out = A.add_noise_batch(c, co, z, no, snr).cpu().numpy()
    outf = A.add_noise_batch(c, co, z, no, snr, out_dtype=torch.float32).cpu().numpy()
    for i, n in enumerate(CASES):
        np.testing.assert_array_equal(out[co[i]:co[i + 1]], golden[f"{n}_mixed"], err_msg=n)
    np.testing.assert_array_equal(outf, out.astype(np.float32))


@pytest.mark.parametrize("seed", [0, 1])
def test_ragged_batch_equals_oracle(seed):
    import torch
    import avsl_b200 as A
    rng = np.random.default_rng(seed)
    lens = [1, 7, 8, 127, 128, 129, 143, 144, 257, 0, 5000, 16001, 48000, 99999, 31, 272]
    nlens = [3, 1, 100, 128, 50, 129, 1000, 7, 256, 4, 5000, 900, 160000, 33333, 31, 17]
    amps = rng.choice([300.0, 3000.0, 15000.0, 40000.0], size=len(lens))     # 40000: beyond int16 -> rescale
    cl = [(rng.standard_normal(n) * a).astype(np.float32) for n, a in zip(lens, amps)]
    nz = [(rng.standard_normal(n) * 2000).astype(np.float32) for n in nlens]
    snr = list(rng.choice([-10, -5, 0, 2.5, 10, 20], size=len(lens)))
    co = np.cumsum([0] + lens)
    no = np.cumsum([0] + nlens)
    want = ON.add_noise_batch(np.concatenate(cl), co, np.concatenate(nz), no, snr)
    got = A.add_noise_batch(torch.from_numpy(np.concatenate(cl)).cuda(), co,
                            torch.from_numpy(np.concatenate(nz)).cuda(), no, snr).cpu().numpy()
    for i in range(len(lens)):
        np.testing.assert_array_equal(got[co[i]:co[i + 1]], want[co[i]:co[i + 1]], err_msg=f"clip {i} len {lens[i]}")
    assert (np.abs(want.astype(np.int32)) >= 32767).any()       # a rescaled clip is among them


def test_edge_cases():
    import torch
    import avsl_b200 as A
    e = torch.empty(0, dtype=torch.float32, device="cuda")
    assert A.add_noise_batch(e, [0], e, [0], 0).numel() == 0
    c = torch.ones(10, dtype=torch.float32, device="cuda") * 100
    with pytest.raises(ZeroDivisionError):
        A.add_noise_batch(c, [0, 10], e, [0, 0], 0)
    with pytest.raises(ValueError):
        A.add_noise_batch(c.double(), [0, 10], c, [0, 10], 0)
    # samples outside the clips pass through unmixed
    z = torch.full((4,), 50.0, device="cuda")
    out = A.add_noise_batch(c, [2, 8], z, [0, 4], 0).cpu().numpy()
    np.testing.assert_array_equal(out[:2], [100, 100])
    np.testing.assert_array_equal(out[8:], [100, 100])
    np.testing.assert_array_equal(out[2:8], ON.add_noise(np.full(6, 100.0), np.full(4, 50.0), 0))


def test_noise_then_logfbank_chain():
    """process_audio_for_av_hubert, preprocess/audio_process.py:222-230: the mixed int16 signal is what
    logfbank sees; the float32 output of the mix feeds logfbank_batch without a host round trip."""
    import torch
    import avsl_b200 as A
    from oracle import logfbank as OF
    clean, noise, snr = MG.noise_long_case()
    clean, noise = clean[:48000], noise[:20000]
    mixed = ON.add_noise(clean, noise, snr)
    want = OF.extract_logfbank_features(mixed.astype(np.float32), stack_order=4)
    m = A.add_noise_batch(torch.from_numpy(clean).cuda(), [0, len(clean)], torch.from_numpy(noise).cuda(),
                          [0, len(noise)], snr, out_dtype=torch.float32)
    np.testing.assert_array_equal(m.cpu().numpy(), mixed.astype(np.float32))
    feats, _ = A.logfbank_batch(m, [0, len(clean)], stack_order=4, normalize=False)
    np.testing.assert_allclose(feats.cpu().numpy(), want, atol=2e-4, rtol=0)


def test_process_audio_for_av_hubert_chain():
    """preprocess/audio_process.py:199-236 from the loaded waveforms on: the noise draw, the mix, the
    stacked logfbank features and the row normalisation against the oracle chain."""
    import avsl_b200 as A
    from oracle import logfbank as OF
    clean, noise, snr = MG.noise_long_case()
    clean, noise = clean[:32000], noise[:9000]

    class Fixed:
        def __init__(self, v):
            self.v = v

        def random(self):
            return self.v
    want_mix = OF.audio_to_tensor(OF.extract_logfbank_features(ON.add_noise(clean, noise, snr).astype(np.float32),
                                                               stack_order=4), normalize=True)
    want_plain = OF.audio_to_tensor(OF.extract_logfbank_features(clean, stack_order=4), normalize=True)
    got_mix = A.process_audio_for_av_hubert(clean, stack_order=4, add_noise_prob=0.5, noise=noise, noise_snr=snr, rng=Fixed(0.1))
    got_skip = A.process_audio_for_av_hubert(clean, stack_order=4, add_noise_prob=0.5, noise=noise, noise_snr=snr, rng=Fixed(0.9))
    got_off = A.process_audio_for_av_hubert(clean, stack_order=4)
    assert got_mix.dtype == np.float32 and got_mix.shape == want_mix.shape
    np.testing.assert_allclose(got_mix, want_mix, atol=1e-3, rtol=0)
    np.testing.assert_allclose(got_skip, want_plain, atol=1e-3, rtol=0)
    np.testing.assert_array_equal(got_off, got_skip)
    assert np.abs(want_mix - want_plain).max() > 0.05            # the mix does change the features
    assert A.process_audio_for_av_hubert(np.zeros(0, np.float32)) is None      # the reference's catch-all


def test_full_size_batch_two_clips_checked():
    """BASELINE-sized batch (64 x 30 s, noise 10 s each, one packed buffer): every clip is compared
    through a size-independent property -- the mix of clip b must not depend on its neighbours, so a
    batch made of one clip repeated has 64 identical answers -- and two clips against the oracle."""
    import torch
    import avsl_b200 as A
    rng = np.random.default_rng(7)
    L, Ln, B = 480000, 160000, 64
    one = rng.integers(-9000, 9001, size=L).astype(np.float32)
    nz = rng.integers(-2500, 2501, size=Ln).astype(np.float32)
    other = rng.integers(-30000, 30001, size=L).astype(np.float32)          # loud: takes the rescale branch at -10 dB
    clean = np.tile(one, B)
    clean[5 * L:6 * L] = other
    snr = [6.0] * B
    snr[5] = -10.0
    got = A.add_noise_batch(torch.from_numpy(clean).cuda(), np.arange(B + 1) * L, torch.from_numpy(np.tile(nz, B)).cuda(),
                            np.arange(B + 1) * Ln, snr).cpu().numpy().reshape(B, L)
    for b in range(B):
        if b != 5:
            assert np.array_equal(got[b], got[0]), b
    np.testing.assert_array_equal(got[0], ON.add_noise(one, nz, 6.0))
    want5, info = ON.add_noise(other, nz, -10.0, return_info=True)
    assert info["rate"] is not None
    np.testing.assert_array_equal(got[5], want5)


@pytest.mark.parametrize("seconds", [70, 231])
def test_long_clips(seconds):
    """70 s: a depth-14 tree, the deepest local heaps of the cluster kernel (one launch per batch; a
    second, short clip in the same batch is handled by CTA 0 of its cluster alone).  231 s: depth 16
    exceeds the cluster kernel's shared-memory heap, the library falls back to separate leaf /
    combine (global-memory path) / mix launches."""
    import torch
    import avsl_b200 as A
    rng = np.random.default_rng(11)
    L = seconds * 16000 + 13
    clean = rng.integers(-8000, 8001, size=L + 4000).astype(np.float32)
    noise = rng.integers(-2000, 2001, size=50001 + 977).astype(np.float32)
    co, no = [0, L, L + 4000], [0, 50001, 50001 + 977]
    want = ON.add_noise_batch(clean, co, noise, no, [3.0, -2.0])
    got = A.add_noise_batch(torch.from_numpy(clean).cuda(), co, torch.from_numpy(noise).cuda(), no, [3.0, -2.0]).cpu().numpy()
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("max_len", [300, 9000, 40000, 140000])
def test_every_cluster_width(max_len):
    """The cluster width follows the longest clip (1, 2, 4, 8 CTAs per clip); every width sees clips
    from one sample up to the longest, unaligned starts, short and long noise."""
    import torch
    import avsl_b200 as A
    rng = np.random.default_rng(max_len)
    lens = [max_len, 1, 7, 129, 255, 256 * 2, 256 * 4 - 1, 256 * 8 + 3, max_len // 2 + 5, max_len - 1]
    lens = [min(n, max_len) for n in lens]
    nlens = [int(rng.integers(1, 3 * n + 2)) for n in lens]
    co = np.concatenate([[0], np.cumsum(lens)])
    no = np.concatenate([[0], np.cumsum(nlens)])
    clean = rng.integers(-9000, 9001, size=co[-1]).astype(np.float32)
    noise = rng.integers(-3000, 3001, size=no[-1]).astype(np.float32)
    snr = [float(v) for v in rng.integers(-15, 20, size=len(lens))]
    want = ON.add_noise_batch(clean, co, noise, no, snr)
    got = A.add_noise_batch(torch.from_numpy(clean).cuda(), co, torch.from_numpy(noise).cuda(), no, snr).cpu().numpy()
    np.testing.assert_array_equal(got, want)


def test_empty_and_tiny_clips_beside_a_long_one():
    """An 8-wide cluster per clip (the longest clip decides): an empty clip, a 5-sample clip and a
    2047-sample clip (each summed by CTA 0 alone) share the launch with a 140,000-sample one."""
    import torch
    import avsl_b200 as A
    rng = np.random.default_rng(77)
    lens, nlens = [140000, 0, 5, 2047], [9000, 0, 3, 5000]
    co, no = np.concatenate([[0], np.cumsum(lens)]), np.concatenate([[0], np.cumsum(nlens)])
    clean = rng.integers(-9000, 9001, size=co[-1]).astype(np.float32)
    noise = rng.integers(-3000, 3001, size=no[-1]).astype(np.float32)
    snr = [5.0, 0.0, -3.0, 12.0]
    want = ON.add_noise_batch(clean, co, noise, no, snr)
    for dt in (torch.int16, torch.float32):
        got = A.add_noise_batch(torch.from_numpy(clean).cuda(), co, torch.from_numpy(noise).cuda(), no, snr, out_dtype=dt)
        np.testing.assert_array_equal(got.cpu().numpy().astype(np.int16), want)
