"""CPU model of the N-adic arithmetic of csrc/nadic.cuh / nadic_inv.cuh, checked against Python integers.

The device code keeps a residue modulo N^2 as two digits x0 + x1*N and does all work modulo N.  This file restates
the exact sequence of steps the kernels execute (two passes per product, the `~m + 1` start of pass 1, the double
conditional subtraction, the chunk-wise lift of plain operands, the exit from the Montgomery domain, the Hensel
step of the inverse) on Python integers, so that the algebra is pinned independently of the GPU parity tests
(tests/test_l012_gpu.py::test_paillier_nadic_edge_cases, tests/test_gg20_gpu.py)."""
import random

import pytest


class Nadic:
    def __init__(self, n: int, k_limbs: int):
        assert n & 1 and n > 1
        self.n, self.bits = n, 32 * k_limbs
        self.R = 1 << self.bits
        self.n_inv = (-pow(n, -1, self.R)) % self.R
        # nadic_setup_kernel: R^2 by doubling (1, 0), then R, R^3, R^4, R^5 by products
        d = (1, 0)
        for _ in range(2 * self.bits):
            d = self.dig_add(d, d)
        self.rr = [None, None, d]
        self.one = self.mul((1, 0), d, cross2=True)
        cur = d
        for _ in range(3):
            cur = self.mul(cur, d, cross2=True)
            self.rr.append(cur)

    # one pass of K rows: (acc0 + x0*b [+ x1*b2] + q*n) / R with the quotient number q
    def _pass(self, acc0, x0, b, x1=0, b2=0):
        t = acc0 + x0 * b + x1 * b2
        q = (t % self.R) * self.n_inv % self.R
        v = t + q * self.n
        assert v % self.R == 0
        return v >> self.bits, q

    def _reduce_twice(self, v):
        cnt = 0
        for _ in range(2):
            if v >= self.n:
                v -= self.n
                cnt += 1
        assert v < self.n
        return v, cnt

    def mul(self, X, Y, cross2):
        """nadic_mul: X*Y*R^-1 mod N^2.  cross2=False drops X1*Y0 (lift: X1 == 0; squaring: Y = (X0, 2*X1 mod N))"""
        (x0, x1), (y0, y1) = X, Y
        u, m = self._pass(0, x0, y0)
        u, uc = self._reduce_twice(u)
        assert uc <= 1
        acc0 = (self.R - 1 - m) + 1                              # ~m in the even set, 1 entering column 0
        t, _ = self._pass(acc0, x0, y1, x1 if cross2 else 0, y0 if cross2 else 0)
        assert t < 3 * self.n
        t, _ = self._reduce_twice(t)
        return u, (t - (1 - uc)) % self.n

    def sqr(self, X):
        return self.mul(X, (X[0], 2 * X[1] % self.n), cross2=False)

    def dig_add(self, A, B):
        lo = A[0] + B[0]
        c = 1 if lo >= self.n else 0
        return lo - c * self.n, (A[1] + B[1] + c) % self.n

    def dig_sub(self, A, B):
        lo = A[0] - B[0]
        b = 1 if lo < 0 else 0
        return lo + b * self.n, (A[1] - B[1] - b) % self.n

    def lift(self, c: int, width_limbs: int):
        """to_nadic: plain operand of `width_limbs` limbs (any value) -> Montgomery digits of c mod N^2"""
        parts = min(4, (width_limbs * 32 + self.bits - 1) // self.bits)
        X = (0, 0)
        for h in range(parts):
            chunk = (c >> (h * self.bits)) % self.R
            X = self.dig_add(X, self.mul((chunk, 0), self.rr[2 + h], cross2=False))
        return X

    def plain(self, X):
        d0, d1 = self.mul(X, (1, 0), cross2=True)
        return d0 + d1 * self.n

    def inverse(self, c: int, width_limbs: int):
        """nadic_inv_kernel: (c mod N)^-1 mod N, then y0*(2 - c*y0)"""
        X = self.lift(c, width_limbs)
        a = X[0] * pow(self.R, -1, self.n) % self.n              # mont_mul(X0, 1) == c mod N
        assert a == c % self.n
        try:
            y0 = pow(a, -1, self.n)
        except ValueError:
            return None
        Y = self.mul((y0, 0), self.rr[2], cross2=False)
        T = self.mul(X, Y, cross2=True)
        W = self.dig_sub(self.dig_add(self.one, self.one), T)
        return self.plain(self.mul(Y, W, cross2=True))


MODULI = [(3, 64), ((1 << 2048) - 1, 64), (5 ** 800, 64), ((1 << 2047) + 1, 64), (65537, 32), ((1 << 1024) - 105, 32)]


@pytest.mark.parametrize("n,k", MODULI)
def test_nadic_products_and_lift(n, k):
    rng = random.Random(n % 1000003)
    A = Nadic(n, k)
    nn, R = n * n, A.R
    assert A.one == (R % nn % n, R % nn // n)
    for h in range(2, 6):
        v = pow(R, h, nn)
        assert A.rr[h] == (v % n, v // n)
    for _ in range(12):
        x, y = rng.getrandbits(4 * A.bits), rng.getrandbits(2 * A.bits)          # operands up to 4K / 2K limbs, mostly >= N^2
        X, Y = A.lift(x, 4 * k), A.lift(y, 2 * k)
        assert X[0] + X[1] * n == x * R % nn and Y[0] + Y[1] * n == y * R % nn
        assert A.plain(A.mul(X, Y, cross2=True)) == x * y % nn
        assert A.plain(A.sqr(X)) == x * x % nn
        assert A.plain(A.dig_add(X, Y)) == (x + y) % nn
    for v in (0, 1, nn - 1, n, n - 1):
        assert A.plain(A.sqr(A.lift(v, 2 * k))) == v * v % nn


@pytest.mark.parametrize("n,k", MODULI)
def test_nadic_exponentiation_and_hensel_inverse(n, k):
    rng = random.Random(7 + n % 1000)
    A = Nadic(n, k)
    nn = n * n
    for _ in range(3):
        b, e = rng.getrandbits(2 * A.bits), rng.getrandbits(200)
        acc, X = A.one, A.lift(b, 2 * k)
        for bit in bin(e)[2:]:
            acc = A.sqr(acc)
            if bit == "1":
                acc = A.mul(acc, X, cross2=True)
        assert A.plain(acc) == pow(b, e, nn)
    for _ in range(6):
        c = rng.getrandbits(2 * A.bits)
        got = A.inverse(c, 2 * k)
        try:
            want = pow(c, -1, nn)
        except ValueError:
            want = None
        assert got == want
    assert A.inverse(0, 2 * k) is None and A.inverse(n * 5, 2 * k) is None
    assert A.inverse(1, 2 * k) == 1
