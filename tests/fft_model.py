"""numpy model of the spectrum core in ``ramannoodle_b200/csrc/rn_fft.cuh`` / ``rn_spectrum.cu``.

The index arithmetic of the CUDA kernels (in-place DIF/DIT stages on 4096-element tiles, the
strided level passes with their digit-reversed level twiddles, the packed/scaled signals, and the
rank-level decimation of the multi-GPU transform) restated with numpy, butterfly for butterfly.
``tests/test_fft_model.py`` checks it against ``numpy.fft`` and the oracle on the CPU, so the
schedule the kernels implement is pinned without a GPU.  Not used by the product.
"""
from __future__ import annotations

import numpy as np

LOG2E = 12
E = 1 << LOG2E
NT = 256
PER = E // NT


def dft_matrix(radix: int, sgn: int) -> np.ndarray:
    q = np.arange(radix)
    return np.exp(sgn * 2j * np.pi * np.outer(q, q) / radix)


def w_root(num, den):
    """exp(-2 pi i num / den)"""
    return np.exp(-2j * np.pi * (np.asarray(num, dtype=np.float64) / den))


def rev3(x, groups):
    x = np.asarray(x).copy()
    r = np.zeros_like(x)
    for _ in range(groups):
        r = (r << 3) | (x & 7)
        x >>= 3
    return r


def fft_stage(tile, sgn, radix, log2sp, log2b):
    """One in-place stage over a (E,) tile — mirrors rn::fft::fft_stage."""
    lr = radix.bit_length() - 1
    sp = 1 << log2sp
    log2ns = log2sp - log2b + lr
    has_tw = log2sp > log2b
    u = np.arange(E // radix)
    lo = u & (sp - 1)
    pos0 = ((u >> log2sp) << (log2sp + lr)) + lo
    pos = pos0[:, None] + (np.arange(radix)[None, :] << log2sp)
    x = tile[pos]
    if has_tw:
        w1 = w_root((lo >> log2b) << (LOG2E - log2ns), E)
        if sgn > 0:
            w1 = w1.conj()
        powers = w1[:, None] ** np.arange(radix)[None, :]
    if sgn > 0 and has_tw:
        x = x * powers
    x = x @ dft_matrix(radix, sgn).T
    if sgn < 0 and has_tw:
        x = x * powers
    tile[pos] = x


def tile_pass(x, hperm=None):
    """tile_kernel on every contiguous 4096-element tile of x (in place): 4096-point DIF, * hperm,
    4096-point DIT.  hperm=None: forward only."""
    for t0 in range(0, x.size, E):
        tile = x[t0:t0 + E]
        for log2sp in (9, 6, 3, 0):
            fft_stage(tile, -1, 8, log2sp, 0)
        if hperm is None:
            continue
        tile *= hperm[t0:t0 + E]
        for log2sp in (0, 3, 6, 9):
            fft_stage(tile, +1, 8, log2sp, 0)


def level_pass(x, sgn, log2lsub, log2r, log2l):
    """level_kernel<sgn> over every sub-array of length 2^log2lsub of x (in place)."""
    log2b = LOG2E - log2r
    log2s = log2lsub - log2r
    b = 1 << log2b
    a8, r1 = divmod(log2r, 3)
    tw_shift = log2l - log2lsub
    big = 1 << log2l
    nsub = x.size >> log2lsub
    for sub in range(nsub):
        for chunk in range(1 << (log2s - log2b)):
            elem_base = (sub << log2lsub) + (chunk << log2b)
            pos = np.arange(E)
            gidx = elem_base + ((pos >> log2b) << log2s) + (pos & (b - 1))
            tile = x[gidx].copy()
            stages = [(8, LOG2E - 3 * (i + 1)) for i in range(a8)]
            if r1:
                stages.append((1 << r1, log2b))
            radix_last, _ = stages[-1]
            lr = radix_last.bit_length() - 1
            groups = (log2r - lr) // 3
            u = np.arange(E // radix_last)
            col = u & (b - 1)
            rb = u >> log2b
            pos0 = (rb << (log2b + lr)) + col
            j = (chunk << log2b) + col
            krest = rev3(rb, groups)
            wb = w_root(((krest * j) << tw_shift) % big, big)
            ws = w_root(((j << (log2r - lr)) << tw_shift) % big, big)
            factors = wb[:, None] * ws[:, None] ** np.arange(radix_last)[None, :]
            lpos = pos0[:, None] + (np.arange(radix_last)[None, :] << log2b)
            if sgn < 0:
                for radix, log2sp in stages[:-1]:
                    fft_stage(tile, -1, radix, log2sp, log2b)
                y = tile[lpos] @ dft_matrix(radix_last, -1).T
                tile[lpos] = y * factors
            else:
                y = tile[lpos] * factors.conj()
                tile[lpos] = y @ dft_matrix(radix_last, +1).T
                for radix, log2sp in reversed(stages[:-1]):
                    fft_stage(tile, +1, radix, log2sp, log2b)
            x[gidx] = tile


def plan_levels(log2lh: int):
    """Strided levels of a local transform of length 2^log2lh (>= 4096): list of log2 R."""
    q = log2lh - LOG2E
    if q < 0:
        raise ValueError("local transform shorter than one tile")
    if q == 0:
        return []
    if q <= 10:
        return [q]
    first = (q + 1) // 2
    return [first, q - first]


def forward(x, log2l):
    """In-place forward transform of x (length 2^log2lh); tables are those of W_{2^log2l}."""
    log2lh = x.size.bit_length() - 1
    log2lsub = log2lh
    for log2r in plan_levels(log2lh):
        level_pass(x, -1, log2lsub, log2r, log2l)
        log2lsub -= log2r
    tile_pass(x)


def convolve(x, hperm, log2l):
    """y = IFFT(FFT(x) * H) (unnormalised) in place, H given in the transform's own order."""
    log2lh = x.size.bit_length() - 1
    levels = plan_levels(log2lh)
    log2lsub = log2lh
    subs = []
    for log2r in levels:
        level_pass(x, -1, log2lsub, log2r, log2l)
        subs.append(log2lsub)
        log2lsub -= log2r
    tile_pass(x, hperm)
    for log2r, lsub in zip(reversed(levels), reversed(subs)):
        level_pass(x, +1, lsub, log2r, log2l)


def frequency_of_position(log2lh: int) -> np.ndarray:
    """k[p]: which frequency bin the forward transform leaves at position p."""
    pos = np.arange(1 << log2lh)
    levels = plan_levels(log2lh)
    shift = log2lh
    k = np.zeros_like(pos)
    weight = 1
    for log2r in levels:
        shift -= log2r
        row = (pos >> shift) & ((1 << log2r) - 1)
        a8, r1 = divmod(log2r, 3)
        low = row & ((1 << r1) - 1)
        k_level = rev3(row >> r1, a8) + (low << (3 * a8))
        k = k + weight * k_level
        weight <<= log2r
    k = k + weight * rev3(pos & (E - 1), 4)
    return k


def filter_layout(log2lh: int) -> np.ndarray:
    """Where the kernels STORE position p of the forward transform (filter spectrum, rn_debug_fft_forward):
    inside every 4096-element tile the eight outputs q of butterfly u (position 8u + q) sit at q * 512 + u,
    so that consecutive threads read consecutive addresses."""
    pos = np.arange(1 << log2lh)
    inner = pos & (E - 1)
    return (pos - inner) + ((inner & 7) << 9) + (inner >> 3)


# ---- chirp-z on top ---------------------------------------------------------------------------
SQ5, SQ525, SQ175, SQ21 = np.sqrt(5.0), np.sqrt(5.25), np.sqrt(1.75), np.sqrt(21.0)


def chirp(n, m_len):
    """exp(-i pi n^2 / M) with n^2 mod 2M formed exactly."""
    n = np.asarray(n, dtype=object)
    phase = np.array([(int(v) * int(v)) % (2 * m_len) for v in n.ravel()], dtype=np.float64).reshape(n.shape)
    return np.exp(-1j * np.pi * phase / m_len)


def packed_signals(alpha):
    """The three complex sequences of measure(): np.diff of the series, six weighted real signals
    packed pairwise, so that  I[k] = (P[k] + P[M-k]) / 4 + E/2  with P = sum_p |FFT_M(z_p)|^2."""
    d = np.diff(alpha, axis=0)
    xx, yy, zz = d[:, 0, 0], d[:, 1, 1], d[:, 2, 2]
    xy, yz, xz = d[:, 0, 1], d[:, 1, 2], d[:, 0, 2]
    return np.stack([SQ5 * (xx + yy + zz) + 1j * SQ525 * (xx - zz),
                     SQ175 * (xx - 2 * yy + zz) + 1j * SQ21 * xy,
                     SQ21 * yz + 1j * SQ21 * xz])


def plan_length(num_frames: int, world: int = 1):
    m_len = num_frames - 1
    log2l = 3
    while (1 << log2l) < max(2 * m_len - 1, E * world):
        log2l += 1
    return m_len, log2l


def filter_sequence(m_len, log2l, world, rank):
    """h_r[n'] of rank `rank`: the circular chirp filter decimated by output residue."""
    big = 1 << log2l
    lh = big // world
    n1 = np.arange(lh)
    acc = np.zeros(lh, dtype=np.complex128)
    for q in range(world):
        idx = q * lh + n1
        m = np.where(idx < m_len, idx, np.where(big - idx < m_len, big - idx, -1))
        h = np.where(m >= 0, chirp(np.maximum(m, 0), m_len).conj(), 0.0)
        acc += h * w_root((q * rank) % world, world)
    return acc * w_root((rank * n1) % big, big)


def md_intensities(alpha, world: int = 1):
    """Uncorrected polycrystalline intensities (bins 1 .. ceil(M/2)-1) the way the kernels compute
    them, with the transform shared by `world` emulated ranks."""
    num_frames = alpha.shape[0]
    m_len, log2l = plan_length(num_frames, world)
    big = 1 << log2l
    lh = big // world
    z = packed_signals(alpha)
    a = z * chirp(np.arange(m_len), m_len)[None, :]
    energy = float((np.abs(z) ** 2).sum())
    padded = np.zeros((3, big), dtype=np.complex128)
    padded[:, :m_len] = a
    hperm = []
    for rank in range(world):
        h = filter_sequence(m_len, log2l, world, rank)
        forward(h, log2l)
        hperm.append(h)
    n1 = np.arange(lh)
    power = np.zeros(m_len)
    for seq in range(3):
        blocks = padded[seq].reshape(world, lh)  # [q][n']
        zr = []
        for rank in range(world):
            b = (blocks * w_root((np.arange(world) * rank) % world, world)[:, None]).sum(axis=0)
            b = b * w_root((rank * n1) % big, big)
            convolve(b, hperm[rank], log2l)
            zr.append(b)
        zr = np.array(zr)  # [r][m']
        for q in range(world):
            m = q * lh + n1
            keep = m < m_len
            if not keep.any():
                continue
            y = (zr * w_root((np.arange(world)[:, None] * n1[None, :]) % big, big).conj()
                 * w_root((np.arange(world) * q) % world, world).conj()[:, None]).sum(axis=0)
            power[m[keep]] += np.abs(y[keep]) ** 2
    points = (m_len + 1) // 2 - 1
    k = np.arange(1, points + 1)
    return (power[k] + power[m_len - k]) * (0.25 / (float(big) * float(big))) + 0.5 * energy


def mirror_half_slots(log2w: int) -> int:
    return (1 << (log2w - 1)) + 64


def mirror_arcs(owner: int, c: int, log2lh: int, log2w: int):
    """csrc/rn_fft.cuh: mirror_arcs — 32-aligned origin of the owner's ascending half, top (31 mod 32) of its
    descending half."""
    lh = 1 << log2lh
    a0 = (owner << log2w) + (c & 1)
    return (((c + a0) >> 1) & (lh - 1)) & ~31, (((2 * lh + c - a0) >> 1) & (lh - 1)) | 31


def mirror_owner(m1: int, c: int, log2lh: int, log2w: int, world: int):
    """csrc/rn_fft.cuh: mirror_owner — (owner rank, slot in the owner's slice of 2 * mirror_half_slots slots)."""
    lh = 1 << log2lh
    u = 2 * m1 - c
    if u < 0:
        u += 2 * lh
    far_side = u > lh
    a = 2 * lh - u if far_side else u
    owner = min(a >> log2w, world - 1)
    near_origin, far_top = mirror_arcs(owner, c, log2lh, log2w)
    if far_side:
        return owner, mirror_half_slots(log2w) + ((far_top - m1) & (lh - 1))
    return owner, (m1 - near_origin) & (lh - 1)


def final_pairs(rank: int, num_bins: int, log2lh: int, world: int):
    """csrc/rn_spectrum.cu: final_dist_kernel — for every pair slot of `rank`: the two residues it reads and
    the bins k it finishes (each with the (q, side) of P[k] and of P[M-k])."""
    lh = 1 << log2lh
    log2w = log2lh - (world.bit_length() - 1)
    w, half = 1 << log2w, 1 << (log2w - 1)
    c = num_bins & (lh - 1)
    points = (num_bins + 1) // 2 - 1
    out = []
    for local in range(half + 1):
        a = (rank << log2w) + 2 * local + (c & 1)
        if not (local < half or (local == half and rank == world - 1 and a == lh)):
            continue
        m_near = ((a + c) >> 1) & (lh - 1)
        m_far = ((2 * lh + c - a) >> 1) & (lh - 1)
        near_origin, far_top = mirror_arcs(rank, c, log2lh, log2w)
        near_slot = (m_near - near_origin) & (lh - 1)
        far_slot = mirror_half_slots(log2w) + ((far_top - m_far) & (lh - 1))
        bins = []
        for q in range(world // 2):
            m = q * lh + m_near
            if m < 1 or m >= num_bins:
                continue
            m2 = num_bins - m
            k = min(m, m2)
            if k > points:
                continue
            bins.append((k, m, m2, m2 >> log2lh))
        out.append((near_slot, far_slot, m_near, m_far, bins))
    return out
