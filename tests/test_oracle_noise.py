"""SNR noise mixing: the oracle against the golden outputs of the reference's own add_noise
(tests/golden/noise_golden.npz, made by tests/golden/make_golden.py::noise_mix), and the spelled-out
pairwise sum against numpy's."""
import hashlib
import os
import sys

import numpy as np
import pytest

from oracle import noise as ON

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as MG  # noqa: E402  (seeded input generators only)

CASES = ["tile", "cut", "equal", "multiple", "tiny", "n130", "n8", "unit_range", "clip_pos", "clip_neg"]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "noise_golden.npz"))


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 127, 128, 129, 136, 143, 144, 1000, 4097, 70001])
def test_pairwise_sum_is_numpys(n):
    x = (np.random.default_rng(n).standard_normal(n) * 3000).astype(np.float32)
    sq = np.square(x)
    assert ON.pairwise_sum(sq) == np.sum(sq)
    if n:
        assert ON.rms(x) == np.sqrt(np.mean(sq, axis=-1))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_output(golden, name):
    clean = golden[f"{name}_clean"].astype(np.float32)
    noise = golden[f"{name}_noise"].astype(np.float32)
    out, info = ON.add_noise(clean, noise, float(golden[f"{name}_snr"]), return_info=True)
    assert out.dtype == np.int16
    np.testing.assert_array_equal(out, golden[f"{name}_mixed"])
    if name == "clip_pos":
        assert info["rate"] is not None and info["max"] >= abs(info["min"])
    if name == "clip_neg":
        assert info["rate"] is not None and info["max"] < abs(info["min"])
    if name == "unit_range":
        assert not out.any()          # the reference's int16 cast zeroes [-1, 1] waveforms


def test_golden_inputs_are_reproducible(golden):
    for name, (clean, noise, snr) in MG.noise_cases().items():
        np.testing.assert_array_equal(clean, golden[f"{name}_clean"].astype(np.float32))
        np.testing.assert_array_equal(noise, golden[f"{name}_noise"].astype(np.float32))


def test_oracle_long_clip(golden):
    clean, noise, snr = MG.noise_long_case()
    out = ON.add_noise(clean, noise, snr)
    np.testing.assert_array_equal(out[golden["long_sel"]], golden["long_mixed_sel"])
    assert hashlib.sha256(out.tobytes()).digest() == golden["long_sha256"].tobytes()


def test_batch_is_per_clip(golden):
    names = ["tile", "tiny", "clip_neg"]
    cl = [golden[f"{n}_clean"].astype(np.float32) for n in names]
    nz = [golden[f"{n}_noise"].astype(np.float32) for n in names]
    co = np.cumsum([0] + [len(c) for c in cl])
    no = np.cumsum([0] + [len(c) for c in nz])
    out = ON.add_noise_batch(np.concatenate(cl), co, np.concatenate(nz), no, [float(golden[f"{n}_snr"]) for n in names])
    for i, n in enumerate(names):
        np.testing.assert_array_equal(out[co[i]:co[i + 1]], golden[f"{n}_mixed"])


def test_noise_plan_host_logic():
    """The shim's bookkeeping (no GPU needed): SNR divisors as the reference's Python expression gives
    them, offsets validation, the reference's ZeroDivisionError for an empty noise clip, workspace
    sized by the library."""
    import torch
    from avsl_b200 import audio as AA
    plan = AA.NoisePlan([0, 10, 10, 500000], [0, 4, 4, 9], [10, 0, -5.5], "cpu")
    assert plan.B == 3 and plan.max_len == 499990 and plan.d_ratio.dtype == torch.float32
    np.testing.assert_array_equal(plan.d_ratio.numpy(), np.array([ON.snr_ratio(10), ON.snr_ratio(0), ON.snr_ratio(-5.5)]))
    assert AA.snr_ratio(20) == np.float32(10.0) and AA.snr_ratio(0) == np.float32(1.0)
    assert plan.ws.numel() >= 3 * 2 * (2 << 13) * 4                      # depth 13 for ~31 s
    assert AA.NoisePlan([0, 5], [0, 2], 3, "cpu").d_ratio.shape == (1,)     # scalar SNR is broadcast
    with pytest.raises(ZeroDivisionError):
        AA.NoisePlan([0, 10], [0, 0], 0, "cpu")
    with pytest.raises(ValueError):
        AA.NoisePlan([0, 10, 5], [0, 1, 2], 0, "cpu")
    with pytest.raises(ValueError):
        AA.NoisePlan([0, 10], [0, 1, 2], 0, "cpu")
    with pytest.raises(ValueError):
        AA.NoisePlan([0, 10], [0, 1], 0, "cpu", out_dtype=torch.float64)
    with pytest.raises(ValueError):
        AA.NoisePlan([0, 1 << 40], [0, 1], 0, "cpu")                     # longer than the library supports


@pytest.mark.parametrize("cs", [2, 4, 8])
@pytest.mark.parametrize("n", [2048, 2049, 4099, 32768, 131072 + 13, 480000, 1120013])
def test_cluster_split_of_the_pairwise_tree(cs, n):
    """What noise_cluster_kernel relies on: CTA `rank` of a cs-wide cluster owns the subtree under heap
    node cs + rank of numpy's pairwise tree -- a contiguous stretch whose bounds come from halving at
    (len >> 1) & ~7 -- and adding the cs subtree sums pairwise, level by level, reproduces
    np.add.reduce bit for bit (n >= 256 * cs: every node above that depth is an inner node)."""
    if n < 256 * cs:
        pytest.skip("short clips are summed by one CTA from the root")
    sq = np.square((np.random.default_rng(n + cs).standard_normal(n) * 3000).astype(np.float32))
    parts = [(0, n)]
    while len(parts) < cs:
        nxt = []
        for off, ln in parts:
            assert ln > 128
            half = (ln >> 1) & ~7
            nxt += [(off, half), (off + half, ln - half)]
        parts = nxt
    sums = [np.sum(sq[off:off + ln]) for off, ln in parts]      # each subtree is itself numpy's tree of its stretch
    while len(sums) > 1:
        sums = [np.float32(sums[i] + sums[i + 1]) for i in range(0, len(sums), 2)]
    assert sums[0] == np.sum(sq)
