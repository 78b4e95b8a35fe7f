"""Oracle: SNR noise mixing (CPU, numpy).  TEST INFRASTRUCTURE ONLY.

Follows ``add_noise(clean_wav, noise_wav, snr)``, preprocess/audio_process.py:110-150 (called from
``process_audio_for_av_hubert`` :222-224 with ``noise_snr`` a Python number).  PARITY PINNED: the
reference's own function was run in the build container on seeded inputs
(tests/golden/make_golden.py::noise_mix -> tests/golden/noise_golden.npz) and this restatement
reproduces every one of those int16 outputs bit for bit (tests/test_oracle_noise.py).

What the reference computes, with the dtypes numpy 2 gives each step (all float32 unless noted):

  1. both waveforms -> float32                                                     (:122-123)
  2. clean_rms = sqrt(mean(clean^2)); the mean is numpy's PAIRWISE float32 sum (spelled out in
     ``pairwise_sum`` below, because the GPU has to add in the same order to match to the bit),
     divided by n in float64 and rounded back to float32                           (:125)
  3. the noise is repeated ceil(Lc / Ln) times when shorter than the clean signal and cut to Lc
     when longer: noise'[i] = noise[i mod Ln], i < Lc                              (:127-132)
  4. noise_rms over noise' (not over the original noise)                           (:134)
  5. gain = (clean_rms / float32(10 ** (snr / 20))) / noise_rms                    (:135-136)
  6. mixed = clean + noise' * gain   (a rounded product, then a rounded sum)       (:136-137)
  7. if max(mixed) > 32767 or min(mixed) < -32768:
         rate = 32767 / max if max >= |min| else -32768 / min;  mixed *= rate      (:140-147)
  8. int16 by truncation toward zero                                               (:149)

Not pinned (the reference has no defined behaviour): NaN / inf samples, a silent noise clip
(noise_rms = 0 -> inf / nan), values outside int16 after step 7.
"""
import numpy as np

F = np.float32


def pairwise_sum(a: np.ndarray) -> np.float32:
    """numpy's float32 add.reduce over a contiguous 1-D array (numpy/_core/src/umath
    ``pairwise_sum``): blocks of at most 128 elements are summed in 8 interleaved accumulators that
    are then combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), the leftover (< 8) elements are added
    one by one; longer arrays are split at n/2 rounded down to a multiple of 8 and the two halves
    added.  Checked against ``np.sum`` itself in tests/test_oracle_noise.py."""
    a = np.ascontiguousarray(a, dtype=F)
    n = a.shape[0]
    if n < 8:
        r = F(0.0)
        for x in a:
            r = F(r + x)
        return r
    if n <= 128:
        body = n - n % 8
        r = a[:body].reshape(-1, 8)
        acc = r[0].copy()
        for row in r[1:]:
            acc = (acc + row).astype(F)
        p = (acc[0::2] + acc[1::2]).astype(F)          # r0+r1, r2+r3, r4+r5, r6+r7
        q = (p[0::2] + p[1::2]).astype(F)
        res = F(q[0] + q[1])
        for x in a[body:]:
            res = F(res + x)
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return F(pairwise_sum(a[:n2]) + pairwise_sum(a[n2:]))


def rms(x: np.ndarray) -> np.float32:
    sq = (x * x).astype(F)
    mean = F(np.float64(pairwise_sum(sq)) / np.float64(len(x)))
    return F(np.sqrt(mean))


def snr_ratio(snr) -> np.float32:
    """float32(10 ** (snr / 20)) -- the Python-float expression of :135 as numpy 2 applies it to a
    float32 scalar (NEP 50: the Python float adopts the array scalar's type)."""
    return F(10 ** (snr / 20))


def add_noise(clean_wav, noise_wav, snr, return_info: bool = False):
    clean = np.asarray(clean_wav).astype(F)
    noise = np.asarray(noise_wav).astype(F)
    n = len(clean)
    tiled = noise[np.arange(n) % len(noise)]
    clean_rms = rms(clean)
    noise_rms = rms(tiled)
    gain = F(F(clean_rms / snr_ratio(snr)) / noise_rms)
    mixed = (clean + (tiled * gain).astype(F)).astype(F)
    hi, lo = mixed.max(), mixed.min()
    rate = None
    if hi > 32767 or lo < -32768:
        rate = F(F(32767) / hi) if hi >= abs(lo) else F(F(-32768) / lo)
        mixed = (mixed * rate).astype(F)
    out = np.trunc(mixed).astype(np.int16)
    if return_info:
        return out, {"clean_rms": clean_rms, "noise_rms": noise_rms, "gain": gain, "max": hi, "min": lo, "rate": rate}
    return out


def add_noise_batch(clean, clean_offsets, noise, noise_offsets, snrs):
    """Packed batch: clip b = clean[co[b]:co[b+1]] mixed with noise[no[b]:no[b+1]] at snrs[b] dB."""
    out = np.zeros(len(clean), np.int16)
    for b in range(len(clean_offsets) - 1):
        c0, c1 = int(clean_offsets[b]), int(clean_offsets[b + 1])
        n0, n1 = int(noise_offsets[b]), int(noise_offsets[b + 1])
        if c1 > c0:
            out[c0:c1] = add_noise(clean[c0:c1], noise[n0:n1], snrs[b])
    return out
