"""CPU checks of the schedule the spectrum kernels implement (tests/fft_model.py restates the index
arithmetic of csrc/rn_fft.cuh and csrc/rn_spectrum.cu in numpy): the in-place DIF transform against
numpy.fft through its digit-reversal map, the convolution (forward, filter, mirrored inverse), and the
whole measure — packed weighted signals, chirp-z, rank-level decimation — against the oracle."""
import numpy as np
import pytest

import fft_model as fm
from oracle import numpy_port as ora


@pytest.mark.parametrize("log2lh", [12, 13, 14, 15, 16, 18])
def test_forward_is_a_permuted_fft(log2lh):
    rng = np.random.default_rng(log2lh)
    n = 1 << log2lh
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    y = x.copy()
    fm.forward(y, log2lh)
    ref = np.fft.fft(x)
    perm = fm.frequency_of_position(log2lh)
    assert sorted(perm.tolist()) == list(range(n))
    assert np.abs(y - ref[perm]).max() / np.abs(ref).max() < 1e-13


def test_filter_layout_is_a_tilewise_permutation():
    layout = fm.filter_layout(14)
    assert sorted(layout.tolist()) == list(range(1 << 14))
    assert np.array_equal(layout // fm.E, np.arange(1 << 14) // fm.E)
    assert layout[8 * 5 + 3] == 3 * 512 + 5 and layout[fm.E + 8 * 511 + 7] == fm.E + 7 * 512 + 511


def test_two_level_plan_forward():
    log2lh = 23  # 2^11 rows above the tiles -> levels of 2^6 and 2^5
    assert fm.plan_levels(log2lh) == [6, 5]
    rng = np.random.default_rng(0)
    n = 1 << log2lh
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    y = x.copy()
    fm.forward(y, log2lh)
    ref = np.fft.fft(x)
    assert np.abs(y - ref[fm.frequency_of_position(log2lh)]).max() / np.abs(ref).max() < 1e-13


@pytest.mark.parametrize("log2lh", [12, 14, 16])
def test_convolution(log2lh):
    rng = np.random.default_rng(100 + log2lh)
    n = 1 << log2lh
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    h = rng.normal(size=n) + 1j * rng.normal(size=n)
    hperm = h.copy()
    fm.forward(hperm, log2lh)
    y = x.copy()
    fm.convolve(y, hperm, log2lh)
    ref = np.fft.ifft(np.fft.fft(x) * np.fft.fft(h)) * n
    assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-13


@pytest.mark.parametrize("frames,world", [(41, 1), (41, 8), (1000, 2), (4097, 4), (5000, 1), (5000, 8), (20_001, 2)])
def test_measure_model_matches_oracle(frames, world):
    rng = np.random.default_rng(frames)
    steps = np.arange(frames)[:, None, None]
    alpha = (6.0 * np.eye(3)[None] + 0.05 * np.sin(0.013 * steps + rng.uniform(0, 6, (1, 3, 3)))
             + 0.01 * rng.normal(size=(frames, 3, 3)))
    _, ref = ora.md_measure(alpha, 1.0)
    got = fm.md_intensities(alpha, world)
    assert np.abs(got / ref - 1).max() < 1e-11


@pytest.mark.parametrize("num_bins, log2lh, world", [(300, 10, 2), (513, 10, 2), (1100, 10, 4), (2047, 10, 4),
                                                     (2048, 10, 4), (4095, 10, 8), (4096, 10, 8), (3330, 10, 8),
                                                     (8999, 12, 8), (8000, 12, 4)])
def test_mirror_ownership_of_shared_transform(num_bins, log2lh, world):
    """Index logic of the shared transform's last two steps (store_out's mirror_owner and the pair
    enumeration of final_dist_kernel), restated in tests/fft_model.py: every residue has exactly one
    (owner, slot); a residue and its mirror image share the owner; the owners' pair lists finish every
    bin k = 1..points exactly once (self-paired residues: twice, with the same operands)."""
    lh = 1 << log2lh
    assert num_bins <= (world // 2) * lh
    log2w = log2lh - (world.bit_length() - 1)
    w = 1 << log2w
    c = num_bins & (lh - 1)
    seen = {}
    for m1 in range(lh):
        owner, slot = fm.mirror_owner(m1, c, log2lh, log2w, world)
        assert 0 <= owner < world and 0 <= slot < 2 * fm.mirror_half_slots(log2w)
        far = slot >= fm.mirror_half_slots(log2w)
        # aligned runs of 32 residues land in one aligned block of 32 slots, ascending or descending
        assert (slot - fm.mirror_half_slots(log2w) if far else slot) % 32 == ((31 - m1 % 32) if far else m1 % 32)
        assert (owner, slot) not in seen
        seen[(owner, slot)] = m1
        mirror = (c - m1) % lh
        assert fm.mirror_owner(mirror, c, log2lh, log2w, world)[0] == owner
    points = (num_bins + 1) // 2 - 1
    finished = {}
    for rank in range(world):
        for near_slot, far_slot, m_near, m_far, bins in fm.final_pairs(rank, num_bins, log2lh, world):
            assert seen[(rank, near_slot)] == m_near
            if m_far != m_near:
                assert seen[(rank, far_slot)] == m_far
            for k, m, m2, q2 in bins:
                assert m % lh == m_near and m2 % lh == m_far and m2 // lh == q2 < world // 2
                assert {m, m2} == {k, num_bins - k}
                if k in finished:
                    assert m_near == m_far  # a self-paired residue meets its pair from both ends
                finished[k] = finished.get(k, 0) + 1
    assert sorted(finished) == list(range(1, points + 1))
