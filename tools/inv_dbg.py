import sys, random, math
sys.path.insert(0, '/root/repo')
import __graft_entry__ as e
from tests.golden import fixtures
pkg = e.load_package(); eng = pkg.Engine(0)
ks = fixtures.load_keyset()
N = ks[0].dk.p * ks[0].dk.q
rng = random.Random(5)
bits = 2048
def check(name, a, mods):
    try:
        got = eng.mod_inv(a, mods, bits)
        ok = all((g == (pow(x, -1, m) if math.gcd(x, m) == 1 else None)) for g, x, m in zip(got, a, mods))
        print(name, "ok" if ok else "WRONG", flush=True)
    except Exception as ex:
        print(name, "EXC", str(ex)[-60:], flush=True); sys.exit(0)
n = 16
check("all-invertible-N", [rng.getrandbits(2040) for _ in range(n)], [N] * n)
check("all-noninv", [ks[0].dk.p * rng.getrandbits(900) for _ in range(n)], [N] * n)
a = [rng.getrandbits(2040) for _ in range(n)]; a[3] = ks[0].dk.p * 77
check("one-noninv", a, [N] * n)
a = [rng.getrandbits(2040) for _ in range(n)]; a[3] = 0
check("one-zero", a, [N] * n)
mods = [rng.getrandbits(bits) | 1 | (1 << (bits - 1)) for _ in range(n)]
check("random-moduli", [rng.getrandbits(bits) for _ in range(n)], mods)
mods[0] = 3
check("with-3", [rng.getrandbits(bits) for _ in range(n)], mods)
