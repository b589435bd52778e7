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
    fr_bound = max(min_fr, max_fr) of a PlaceCells / GridCells population without OU noise: with RIAB_THIN_SPIKES set in the
    environment the library uses the (experimental, measured slower) thinned stream when dt * fr_bound * (1 + 2^-10) <= 1/8
    (expected_spikes_thin); otherwise -- the default, and always for every other population -- the dense one:
        spike <=> m < fma(rate, dt*65536, -v)      (riab_b200.cu: spike_ballots)
    One Philox4x32-7 call per (agent pair gid>>1, 4-cell group); agent half gid&1 takes words 2h, 2h+1 as four
    16-bit integers m; all eight share the dither v = ((r0^r2)>>8) * 2^-24.  The float32 fma is mirrored
    in float64: the 48-bit product and the 24-bit dither add exactly, so the sum rounds to float32 once."""
    import os
    if fr_bound is not None and os.environ.get("RIAB_THIN_SPIKES") and thin_tables(dt, fr_bound) is not None:
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


STREAM_THIN_FIRST, STREAM_THIN_CHAIN = 4, 8


def thin_tables(dt, fr_bound):
    """Thresholds of the thinned spike stream (riab_b200.cu: make_out).  None when the dense stream is used."""
    bound = float(fr_bound) * (1.0 + 1.0 / 1024.0)
    p = float(dt) * bound
    if not (fr_bound >= 0.0 and p <= 0.125):
        return None
    q8 = 1.0 - p
    q8 *= q8; q8 *= q8; q8 *= q8
    t16 = min(np.ceil(65536.0 * (1.0 - q8)), 65535.0)
    q1 = np.sqrt(np.sqrt(np.sqrt(1.0 - t16 / 65536.0)))
    pp = 1.0 - q1
    cum, qq = [], 1.0
    for _ in range(8):
        qq *= q1
        cum.append(1.0 - qq)
    clamp = lambda v: 4294967295 if v >= 4294967295.0 else int(v)
    t = [clamp(np.floor(4294967296.0 * c)) for c in cum]
    tc = [clamp(np.floor(4294967296.0 * (cum[i] / cum[7]))) if cum[7] > 0.0 else 0 for i in range(7)]
    bnd = pp / float(dt) if dt > 0 else 0.0
    c1, c0 = np.float32(bnd / 16777216.0), np.float32(bnd / 33554432.0)
    return int(t16), np.array(t, dtype=np.uint64), np.array(tc, dtype=np.uint64), c1, c0


def expected_spikes_thin(seed, step, agents, fr, dt, fr_bound, pop=0):
    """Mirror of the thinned spike stream (riab_b200.cu: thin_rows / thin_pass): candidates per octet = (agent pair, 4-cell group)
    of 8 slots, slot = 4*(gid&1) + cell&3.
    Level 1: half-word (gid>>1)&7 of Philox7((gid>>4, group), THIN_FIRST) < t16  <=>  the octet holds a candidate.
    Level 2: words of Philox7((gid>>1, group), THIN_CHAIN + n), n = 0, 1, ...: first slot from the conditional table,
    then alternately accept (24-bit uniform * bound < rate) and geometric gap to the next candidate."""
    agents = np.asarray(agents, dtype=np.uint64)
    fr = np.asarray(fr, dtype=np.float32)
    A, n_cells = fr.shape
    t16, t, tc, c1, c0 = thin_tables(dt, fr_bound)
    key = (seed & 0xFFFFFFFF, seed >> 32)
    groups = (n_cells + 3) // 4
    out = np.zeros((A, n_cells), dtype=bool)
    row_of = {int(g): i for i, g in enumerate(agents)}
    pairs = np.unique(agents >> np.uint64(1))
    g = np.arange(groups, dtype=np.uint64)[None, :]
    first = philox4x32(counter((pairs >> np.uint64(3))[:, None], g, step, STREAM_THIN_FIRST, pop), key, rounds=7)   # (P,G,4)
    h = (pairs & np.uint64(7)).astype(np.int64)
    word = np.take_along_axis(first, (h >> 1)[:, None, None].repeat(groups, 1), axis=2)[..., 0]
    half = np.where((h & 1)[:, None] == 1, word >> np.uint32(16), word & np.uint32(0xFFFF))
    cand_p, cand_g = np.nonzero(half < np.uint32(t16))

    def gap(x):
        x = np.uint64(x)
        return 8 if x >= t[7] else int((x >= t[:7]).sum())

    def accept(pair, grp, K, word):
        gid = 2 * int(pair) + (K >> 2)
        cell = 4 * int(grp) + (K & 3)
        if gid not in row_of or cell >= n_cells:
            return
        thr = np.float32(np.float64(int(word) >> 8) * np.float64(c1) + np.float64(c0))
        if thr < fr[row_of[gid], cell]:
            out[row_of[gid], cell] = True

    for pi, gi in zip(cand_p, cand_g):
        pair, n = pairs[pi], 0
        S = philox4x32(counter(np.uint64(pair), np.uint64(gi), step, STREAM_THIN_CHAIN + n, pop), key, rounds=7)
        K = int((np.uint64(S[0]) >= tc).sum())
        accept(pair, gi, K, S[1])
        K += 1 + gap(S[2])
        while K < 8:
            accept(pair, gi, K, S[3])
            n += 1
            S = philox4x32(counter(np.uint64(pair), np.uint64(gi), step, STREAM_THIN_CHAIN + n, pop), key, rounds=7)
            K += 1 + gap(S[0])
            if K >= 8:
                break
            accept(pair, gi, K, S[1])
            K += 1 + gap(S[2])
    return out
