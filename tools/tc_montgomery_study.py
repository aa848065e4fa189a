"""Design study (CPU, numpy): moving the two CONSTANT-operand products of a Montgomery multiplication to the
int8 tensor-core path.  Not part of the product; see DESIGN.md section 8 ("next").

A Montgomery product a*b*R^-1 mod N in separated-operand form is three big products
    T = a * b                   (operands private to the job)                 -> INT32 multiply-add pipe
    m = (T mod R) * N' mod R    (N' = -N^-1 mod R: a per-KEY constant)        -> tensor cores
    U = (T + m * N) / R         (N: a per-KEY constant)                       -> tensor cores
Both constant products are digit convolutions, i.e. Toeplitz(constant) x digit-vector; for the jobs of one
warp that share a key they are one small GEMM with u8 x u8 -> s32 MMAs (mma.sync.m16n8k32 / tcgen05 kind::i8):
rows = jobs, K-dim = the 8-bit digits of the variable operand, columns = output digit positions.
With 4K <= 256 digits every s32 accumulator stays below 256 * 255^2 < 2^24, so the result is exact; the digits
are then carry-normalised back into 32-bit limbs (ALU pipe).  What stays on the IMAD pipe is a*b alone: K^2 of
the 2K^2 wide MACs of the interleaved form the engine runs today.

This file checks the arithmetic of that scheme bit-exactly against Python integers and prints the operation
counts the DESIGN note quotes.  `python tools/tc_montgomery_study.py`
"""
from __future__ import annotations

import random

import numpy as np


def digits8(x: int, n: int) -> np.ndarray:
    return np.frombuffer(x.to_bytes(n, "little"), dtype=np.uint8).astype(np.int64)


def toeplitz(c: int, n_in: int, n_out: int) -> np.ndarray:
    """[n_in, n_out] matrix with M[i, k] = digit_{k-i}(c): row-vector of digits x M = product digits (un-normalised)"""
    d = digits8(c, n_out)
    m = np.zeros((n_in, n_out), dtype=np.int64)
    for i in range(n_in):
        m[i, i:] = d[: n_out - i]
    return m


def normalise(cols: np.ndarray) -> int:
    """sum_k cols[k] * 2^(8k) for un-normalised s32 digit columns"""
    v = 0
    for k, c in enumerate(cols.tolist()):
        v += int(c) << (8 * k)
    return v


def mont_mul_tc(a: np.ndarray, b: np.ndarray, n: int, bits: int):
    """batch of Montgomery products over ONE modulus n; a, b: python-int lists.  Returns (results, max accumulator)."""
    nd = bits // 8
    R = 1 << bits
    n_prime = (-pow(n, -1, R)) % R
    T_np = toeplitz(n_prime, nd, nd)            # low half only: m = T_lo * N' mod R
    T_n = toeplitz(n, nd, 2 * nd)
    out, peak = [], 0
    t_all = [x * y for x, y in zip(a, b)]                                        # IMAD pipe: K^2 wide MACs per job
    lo = np.stack([digits8(t % R, nd) for t in t_all])                          # [jobs, nd] u8
    acc_m = lo @ T_np                                                            # tensor cores: u8 x u8 -> s32
    peak = max(peak, int(acc_m.max()))
    m_all = [normalise(r) % R for r in acc_m]                                    # ALU: carry normalisation, keep the low K limbs
    md = np.stack([digits8(m, nd) for m in m_all])
    acc_u = md @ T_n                                                             # tensor cores
    peak = max(peak, int(acc_u.max()))
    for t, r in zip(t_all, acc_u):
        u = (t + normalise(r)) >> bits                                           # exact: low half cancels
        assert (t + normalise(r)) % R == 0
        out.append(u - n if u >= n else u)
    return out, peak


def main():
    rng = random.Random(0xB200)
    for bits in (1024, 2048):
        n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        R = 1 << bits
        jobs = 8                                                                 # one warp of TPI = 4 lane groups
        a = [rng.randrange(n) for _ in range(jobs)]
        b = [rng.randrange(n) for _ in range(jobs)]
        a[0], b[0] = n - 1, n - 1
        got, peak = mont_mul_tc(a, b, n, bits)
        want = [x * y * pow(R, -1, n) % n for x, y in zip(a, b)]
        assert got == want
        K = bits // 32
        nd = bits // 8
        imad_now = 2 * K * K
        imad_new = K * K
        mma_macs = nd * nd + nd * 2 * nd                                         # u8 MACs per job (dense Toeplitz tiles, both products)
        print(f"{bits}-bit: exact; max s32 accumulator {peak} (< 2^{peak.bit_length()}); per product and job: wide MACs {imad_now} -> {imad_new}, "
              f"u8 tensor MACs {mma_macs} ({mma_macs // 4096} m16n8k32 tiles at full row use), digit columns to normalise {3 * nd}")


if __name__ == "__main__":
    main()
