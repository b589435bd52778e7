"""NumPy Philox4x32 (Salmon et al. SC'11) mirroring ratinabox_b200/csrc/riab_common.cuh:
counter layout, key schedule, uniform / normal conversions.  Test infrastructure."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
STREAM_AGENT_OU, STREAM_CELL_NOISE, STREAM_SPIKES, STREAM_MEASURE = 0, 1, 2, 3


def philox4x32(ctr, key, rounds=10):
    """ctr: (...,4) uint32, key: (2,) ints -> (...,4) uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def counter(agent, sub, step, stream, pop=0):
    agent = np.asarray(agent, dtype=np.uint64)
    sub = np.asarray(sub, dtype=np.uint64)
    c = np.zeros(np.broadcast(agent, sub).shape + (4,), dtype=np.uint32)
    c[..., 0] = (agent & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c[..., 1] = ((sub ^ ((agent >> np.uint64(32)) << np.uint64(24))) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c[..., 2] = np.uint32(step & 0xFFFFFFFF)
    c[..., 3] = np.uint32(((step >> 32) & 0xFFFF) | ((pop & 0xFF) << 16) | (stream << 24))
    return c


def u01_53(hi, lo):
    x = ((hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)) >> np.uint64(11)
    return (x.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def u01_24(x):
    # fma(x>>8, 2^-24, 2^-25): exact in float32 ((x>>8)+0.5 has 25 significant bits only when x>>8 >= 2^23,
    # where the fused result rounds once) -- mirror with float64 arithmetic rounded once to float32
    return ((x >> np.uint32(8)).astype(np.float64) * 2.0 ** -24 + 2.0 ** -25).astype(np.float32)


def agent_normals(seed, step, agents):
    """The two standard normals of Agent.update's OU draws: Box-Muller in float32 on two 32-bit
    uniforms (mirrors riab_common.cuh: agent_normals; equal up to the last float32 ulps of
    logf / sincospif)."""
    r = philox4x32(counter(agents, 0, step, STREAM_AGENT_OU), (seed & 0xFFFFFFFF, seed >> 32))
    f32 = np.float32
    u1 = (r[..., 0].astype(f32).astype(np.float64) * 2.0 ** -32 + 2.0 ** -33).astype(f32)
    u2 = (r[..., 2].astype(f32) * f32(2.0 ** -32)).astype(f32)
    rad = np.sqrt(f32(-2.0) * np.log(u1)).astype(f32)
    ang = (f32(2.0) * u2).astype(np.float64) * np.pi
    return np.stack(((rad * np.cos(ang).astype(f32)).astype(np.float64),
                     (rad * np.sin(ang).astype(f32)).astype(np.float64)), axis=-1)


def expected_spikes(seed, step, agents, fr, dt, pop=0, fr_bound=None):
    """(A, n_cells) bool spikes of one step given the float32 rates `fr` (A, n_cells).
    fr_bound = max(min_fr, max_fr) of a GridCells population without OU noise: the library uses the thinned
    stream when dt * fr_bound * (1 + 2^-10) <= 1/16 (expected_spikes_thin; RIAB_DENSE_SPIKES=1 in the environment turns it
    off); otherwise -- and always for every other population (expected_spikes_of) -- the dense one:
        spike <=> m < fma(rate, dt*65536, -v)      (riab_b200.cu: spike_ballots)
    One Philox4x32-7 call per (agent pair gid>>1, 4-cell group); agent half gid&1 takes words 2h, 2h+1 as four
    16-bit integers m; all eight share the dither v = ((r0^r2)>>8) * 2^-24.  The float32 fma is mirrored
    in float64: the 48-bit product and the 24-bit dither add exactly, so the sum rounds to float32 once."""
    import os
    if fr_bound is not None and not os.environ.get("RIAB_DENSE_SPIKES") and thin_tables(dt, fr_bound) is not None:
        return expected_spikes_thin(seed, step, agents, fr, dt, fr_bound, pop)
    agents = np.asarray(agents, dtype=np.uint64)
    fr = np.asarray(fr, dtype=np.float32)
    n_cells = fr.shape[1]
    groups = (n_cells + 3) // 4
    a = (agents >> np.uint64(1))[:, None]
    g = np.arange(groups, dtype=np.uint64)[None, :]
    r = philox4x32(counter(a, g, step, STREAM_SPIKES, pop), (seed & 0xFFFFFFFF, seed >> 32), rounds=7)   # (A,G,4)
    v = ((r[..., 0] ^ r[..., 2]) >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    h = (agents & np.uint64(1)).astype(np.int64)[:, None]
    w0 = np.where(h == 1, r[..., 2], r[..., 0])
    w1 = np.where(h == 1, r[..., 3], r[..., 1])
    m = np.stack((w0 & np.uint32(0xFFFF), w0 >> np.uint32(16), w1 & np.uint32(0xFFFF), w1 >> np.uint32(16)), axis=-1)
    m = m.astype(np.float32).reshape(len(agents), groups * 4)[:, :n_cells]
    vv = np.repeat(v, 4, axis=1)[:, :n_cells]
    q = np.float64(np.float32(np.float32(dt) * np.float32(65536.0)))
    thr = (fr.astype(np.float64) * q - vv).astype(np.float32)
    return m < thr


def expected_spikes_of(ns, seed, step, agents, fr, dt, pop=0, fr_bound=None):
    """expected_spikes for a ratinabox_b200 population: the library uses the thinned stream for GridCells only (bounded
    rates and a pair loop for which it measured faster; riab_b200.cu: GridPolicy::THIN, launch_tile) -- PlaceCells keep the
    dense stream whatever their rate bound."""
    thin_ok = type(ns).__name__ == "GridCells"
    return expected_spikes(seed, step, agents, fr, dt, pop=pop, fr_bound=fr_bound if thin_ok else None)


STREAM_THIN = 4


def thin_tables(dt, fr_bound):
    """Tables of the thinned spike stream (riab_b200.cu: make_out; the same IEEE operations in the same order).
    None when the dense stream is used.  cdf[k] = floor(2^32 P(Binomial(128, p') <= k)), k < 32."""
    bound = float(fr_bound) * (1.0 + 1.0 / 1024.0)
    p = float(dt) * bound
    if not (fr_bound >= 0.0 and 0.0 < p <= 0.0625):
        return None
    q = 1.0 - p
    r = p / q
    pmf = q
    for _ in range(7):
        pmf = pmf * pmf
    cdf, tab = 0.0, []
    for i in range(32):
        cdf = cdf + pmf
        t = np.floor(4294967296.0 * cdf)
        tab.append(4294967295 if t >= 4294967295.0 else int(t))
        pmf = pmf * (float(128 - i) * r) / float(i + 1)
    return np.array(tab, dtype=np.uint64), np.float32(bound / 1048576.0), np.float32(bound / 2097152.0)


def expected_spikes_thin(seed, step, agents, fr, dt, fr_bound, pop=0):
    """Mirror of the thinned spike stream (riab_b200.cu: thin_block).  Per (row gid, 128-cell block B), calls n = 0, 1, ...
    R_n = Philox7(ctr(gid, B | n << 16, step, THIN, pop)):  K = #{k: R_0[0] >= cdf[k]} candidates; draw d = 4n + j has the
    position (R_n[1] >> 7j) & 127 and the 20-bit uniform x = half-word j of (R_n[2], R_n[3]) << 4 | R_n[1] >> 28; the
    candidates are the first K distinct positions of the draw sequence; a candidate spikes iff fma(x, c1, c0) < rate."""
    agents = np.asarray(agents, dtype=np.uint64)
    fr = np.asarray(fr, dtype=np.float32)
    A, n_cells = fr.shape
    cdf, c1, c0 = thin_tables(dt, fr_bound)
    key = (seed & 0xFFFFFFFF, seed >> 32)
    nb = (n_cells + 127) // 128
    frp = np.zeros((A, nb * 128), dtype=np.float32)
    frp[:, :n_cells] = fr
    frp = frp.reshape(A, nb, 128)
    exists = (np.arange(nb * 128) < n_cells).reshape(1, nb, 128)
    gid = agents[:, None]
    blk = np.arange(nb, dtype=np.uint64)[None, :]
    R = philox4x32(counter(gid, blk, step, STREAM_THIN, pop), key, rounds=7)                       # (A, nb, 4)
    K = (R[..., 0].astype(np.uint64)[..., None] >= cdf[None, None, :]).sum(-1)                     # (A, nb)
    occ = np.zeros((A, nb, 128), dtype=bool)
    out = np.zeros((A, nb, 128), dtype=bool)
    cnt = np.zeros((A, nb), dtype=np.int64)
    ai, bi = np.meshgrid(np.arange(A), np.arange(nb), indexing="ij")
    n = 0
    while True:
        dith = R[..., 1] >> np.uint32(28)
        for j in range(4):
            pos = ((R[..., 1] >> np.uint32(7 * j)) & np.uint32(127)).astype(np.int64)
            take = (cnt < K) & ~occ[ai, bi, pos]
            occ[ai, bi, pos] |= take
            cnt += take
            w = R[..., 2] if j < 2 else R[..., 3]
            m = (w >> np.uint32(16 * (j & 1))) & np.uint32(0xFFFF)
            x = ((m << np.uint32(4)) | dith).astype(np.float64)
            thr = (x * np.float64(c1) + np.float64(c0)).astype(np.float32)      # the fused multiply-add rounds once
            hit = take & exists[0, bi, pos] & (thr < frp[ai, bi, pos])
            out[ai, bi, pos] |= hit
        n += 1
        if not (cnt < K).any() or n >= 256:
            break
        R = philox4x32(counter(gid, blk | np.uint64(n << 16), step, STREAM_THIN, pop), key, rounds=7)
    return out.reshape(A, nb * 128)[:, :n_cells]
