"""Known-answer tests for the Philox4x32-10 used by the CUDA path (NumPy mirror).  CPU only."""
import numpy as np

from philox_np import philox4x32


def test_philox4x32_10_kat():
    # Random123 kat_vectors: philox4x32_10
    cases = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in cases:
        got = philox4x32(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(x) for x in got) == want


def test_thin_tables_are_the_binomial_cdf():
    """thin_tables repeats riab_b200.cu: make_out operation for operation: cdf[k] = floor(2^32 P(Binomial(128, p') <= k)),
    p' = dt * bound * (1 + 2^-10); None (dense stream) above 1/16."""
    from math import comb
    from philox_np import thin_tables
    cdf, c1, c0 = thin_tables(0.01, 1.0)
    p = 0.01 * (1.0 + 1.0 / 1024.0)
    exact = np.cumsum([comb(128, k) * p ** k * (1 - p) ** (128 - k) for k in range(32)])
    assert np.abs(cdf.astype(np.float64) / 2.0 ** 32 - exact).max() < 1e-9
    assert np.all(np.diff(cdf.astype(np.int64)) >= 0) and int(cdf[-1]) == 2 ** 32 - 1
    assert c1 == np.float32((1.0 + 1.0 / 1024.0) / 1048576.0) and c0 == np.float32((1.0 + 1.0 / 1024.0) / 2097152.0)
    assert thin_tables(0.05, 15.0) is None and thin_tables(0.01, 6.3) is None and thin_tables(0.01, 6.2) is not None


def test_thinned_stream_is_bernoulli_dt_rate():
    """The thinned spike stream's mirror (what the GPU tests hold the CUDA path to, bit for bit) is Bernoulli(dt * rate) per
    (agent, cell) (Neurons.py:681-684): totals, per-cell totals and per-rate-band frequencies over 2.6e7 draws, and
    independence of the sharding (rows keyed by global agent id)."""
    from philox_np import expected_spikes
    rs = np.random.RandomState(1)
    A, N, dt = 20000, 260, 0.02
    fr = rs.rand(A, N).astype(np.float32)
    p = dt * fr.astype(np.float64)
    tot, hits = 0, np.zeros(N)
    for s in range(5):
        sp = expected_spikes(7, s, np.arange(A), fr, dt, pop=0, fr_bound=1.0)
        tot += sp.sum()
        hits += sp.sum(0)
    mu, var = 5 * p.sum(), 5 * (p * (1 - p)).sum()
    assert abs(tot - mu) < 5 * np.sqrt(var)
    z = (hits - 5 * p.sum(0)) / np.sqrt(5 * (p * (1 - p)).sum(0))
    assert np.abs(z).max() < 5 and abs(z.std() - 1) < 0.2
    sp = expected_spikes(7, 9, np.arange(A), fr, dt, pop=0, fr_bound=1.0)
    for lo in np.arange(0, 1, 0.25):
        m = (fr >= lo) & (fr < lo + 0.25)
        assert abs(sp[m].sum() - p[m].sum()) < 4.5 * np.sqrt((p[m] * (1 - p[m])).sum())
    part = expected_spikes(7, 9, np.arange(1000, 1500), fr[1000:1500], dt, pop=0, fr_bound=1.0)
    assert np.array_equal(part, sp[1000:1500])
