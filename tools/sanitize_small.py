"""Tiny end-to-end run for compute-sanitizer: one offline session + a few L0/L2 calls."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tests.golden import fixtures
from oracle.sampling import Drbg, sample_unit
pkg = entry.load_package()
from mpecdsa_b200 import gg20
eng = pkg.Engine(0)
keyset = fixtures.load_keyset(0)
ks = gg20.KeySets(eng, [keyset])
rng = Drbg(1, "san")
keys = [keyset[0], keyset[1]]
rnd = [sample_unit(rng, keys, [1, 2], p) for p in range(2)]
res = gg20.offline_batch(eng, ks, [(0, 0, 1)], gg20.pack_randomness(rnd))
print("offline status", list(res.status), flush=True)
r = random.Random(2)
m = [r.getrandbits(2048) | 1 | (1 << 2047) for _ in range(5)]
print("modexp", eng.mod_pow([r.getrandbits(2048) for _ in m], [r.getrandbits(300) for _ in m], m)[1].tolist(), flush=True)
print("modinv", [x is not None for x in eng.mod_inv([r.getrandbits(2048) for _ in m], m)], flush=True)
print("secp", eng.secp_mul(None, [5, 7])[0] is not None, flush=True)
ks.free(); eng.close()
print("done", flush=True)
