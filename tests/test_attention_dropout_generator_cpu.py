"""Integer restatement of the attention-probability dropout generator of csrc/attention.cu (attn_row_key / attn_pair_bits):
keep mask of element (query row, key) = 16 bits of mix32(rowkey ^ (key >> 1) * 0x9E3779B1), rowkey = folded splitmix64 of
(seed, launch tick, row).  Checks what the training semantics need from it: the drop rate, no correlation between adjacent
keys / the two keys of a pair / adjacent rows / keys 8 apart (the two key rows one dK/dV-pass thread owns), and binomial
row / column spreads.  (The GPU tests check that forward, dQ pass and dK/dV pass evaluate the SAME mask.)"""
import numpy as np

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


def row_key(seed, tick, row):
    z = (seed + tick * 0x2545F4914F6CDD1D + row * 0x9E3779B97F4A7C15) & M64          # common.cuh: dropout_bits4
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return (z ^ (z >> 32)) & M32


def pair_bits(rk, kp):
    h = (rk ^ ((kp * 0x9E3779B1) & M32)) & M32
    h = (h * 0x85EBCA6B) & M32
    h ^= h >> 15
    h = (h * 0xC2B2AE35) & M32
    h ^= h >> 16
    return h


def test_generator_statistics():
    R, S, p = 384, 256, 0.1
    thr = int(p * 65536)
    drop = np.zeros((R, S), dtype=np.float64)
    for r in range(R):
        rk = row_key(1234, 7, r)
        for kp in range(S // 2):
            h = pair_bits(rk, kp)
            drop[r, 2 * kp] = (h & 0xFFFF) < thr
            drop[r, 2 * kp + 1] = (h >> 16) < thr
    assert abs(drop.mean() - p) < 3e-3
    m = drop - drop.mean()
    var = m.var()
    for a, b in ((m[:, :-1], m[:, 1:]), (m[:, 0::2], m[:, 1::2]), (m[:-1], m[1:]), (m[:, :-8], m[:, 8:])):
        assert abs((a * b).mean() / var) < 1.5e-2
    assert drop.mean(0).std() < 1.3 * np.sqrt(p * (1 - p) / R)
    assert drop.mean(1).std() < 1.3 * np.sqrt(p * (1 - p) / S)


def test_masks_change_with_the_launch_tick_and_the_seed():
    a = [pair_bits(row_key(5, 1, r), 3) for r in range(64)]
    b = [pair_bits(row_key(5, 2, r), 3) for r in range(64)]
    c = [pair_bits(row_key(6, 1, r), 3) for r in range(64)]
    assert a != b and a != c and len(set(a)) == 64
