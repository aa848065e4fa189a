"""BASELINE.json configs[2] and configs[3] at full size on one GPU, with spot checks against independent implementations.
Prints one JSON line per configuration (kept under profiles/)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry
from oracle import gg20_oracle as o
from tests.golden import fixtures
pkg = entry.load_package()
from mpecdsa_b200 import gg20
eng = pkg.Engine(0)
pkg._bind_l01(eng.lib); gg20._bind_l2(eng.lib)
rng = np.random.default_rng(0xB2000003)

# ---- configs[2]: 1M secp256k1 scalar-mul (random scalars, generator base) + a variable-base run
n = 1 << 20
k = rng.integers(0, 2**32, size=(n, 8), dtype=np.uint32); k[:, 7] &= 0x7FFFFFFF
out = np.zeros((n, 16), dtype=np.uint32)
for rep in range(2):
    t0 = time.perf_counter()
    eng._ck(eng.lib.tecdsa_secp_mul_batch(eng._ctx, None, k.ctypes.data, out.ctypes.data, n, 0), "secp_mul")
    dt_gen = time.perf_counter() - t0
from cryptography.hazmat.primitives.asymmetric import ec
idx = np.linspace(0, n - 1, 256).astype(np.int64)
ok = True
for i in idx:
    ki = int.from_bytes(k[i].tobytes(), "little")
    pub = ec.derive_private_key(ki, ec.SECP256K1()).public_key().public_numbers()
    v = int.from_bytes(out[i].tobytes(), "little")
    ok = ok and (v & ((1 << 256) - 1), v >> 256) == (pub.x, pub.y)
nv = 1 << 18
pts = out[:nv].copy()
out2 = np.zeros((nv, 16), dtype=np.uint32)
t0 = time.perf_counter()
eng._ck(eng.lib.tecdsa_secp_mul_batch(eng._ctx, pts.ctypes.data, k[nv:2 * nv].ctypes.data, out2.ctypes.data, nv, 0), "secp_mul")
dt_var = time.perf_counter() - t0
# (k2 * (k1 G)) == ((k1*k2) G) on a sample
for i in range(0, nv, nv // 64):
    k1 = int.from_bytes(k[i].tobytes(), "little"); k2 = int.from_bytes(k[nv + i].tobytes(), "little")
    v = int.from_bytes(out2[i].tobytes(), "little")
    ok = ok and (v & ((1 << 256) - 1), v >> 256) == o.pt_mul(o.G, k1 * k2 % o.Q)
print(json.dumps({"config": "configs[2]: 1M secp256k1 scalar-mul, generator base", "count": n, "seconds_e2e_host_buffers": dt_gen,
                  "scalar_mul_per_s": n / dt_gen, "variable_base": {"count": nv, "seconds": dt_var, "scalar_mul_per_s": nv / dt_var},
                  "checked_vs_openssl": 256, "ok": bool(ok)}), flush=True)

# ---- configs[3]: 16k MtA range proofs generate + verify
keyset = fixtures.load_keyset(0)
ks = gg20.KeySets(eng, [keyset, fixtures.load_keyset(1)])
m = 16384
ek_row = rng.integers(0, 6, size=m).astype(np.uint32)
st_row = rng.integers(0, 6, size=m).astype(np.uint32)
ns = [lk.dk.p * lk.dk.q for kset in (keyset, fixtures.load_keyset(1)) for lk in kset]
nts = [kset[i].h1_h2_n_tilde_vec[i].N for kset in (keyset, fixtures.load_keyset(1)) for i in range(3)]
def rnd_bits(rows_bits, limbs):
    a = rng.integers(0, 2**32, size=(m, limbs), dtype=np.uint32)
    for i in range(m):
        b = rows_bits[i]; full, rem = divmod(b, 32)
        a[i, full + (1 if rem else 0):] = 0
        if rem: a[i, full] &= (1 << rem) - 1
    a[:, 0] |= 1
    return a
q3b = (o.Q ** 3).bit_length() - 1
a_ = rnd_bits([255] * m, 8)
r_ = rnd_bits([ns[e].bit_length() - 1 for e in ek_row], 64)
al = rnd_bits([q3b] * m, 24); be = rnd_bits([ns[e].bit_length() - 1 for e in ek_row], 64)
ga = rnd_bits([(o.Q ** 3 * nts[s]).bit_length() - 1 for s in st_row], 88); ro = rnd_bits([(o.Q * nts[s]).bit_length() - 1 for s in st_row], 72)
N = pkg.ints_to_limbs(ns, 64)
a64 = np.zeros((m, 64), np.uint32); a64[:, :8] = a_
c = np.zeros((m, 128), np.uint32)
eng._ck(eng.lib.tecdsa_paillier_encrypt_batch(eng._ctx, N.ctypes.data, ek_row.ctypes.data, 6, a64.ctypes.data, r_.ctypes.data, c.ctypes.data, m, 0), "enc")
z = np.zeros((m, 64), np.uint32); e = np.zeros((m, 8), np.uint32); s = np.zeros((m, 64), np.uint32); s1 = np.zeros((m, 28), np.uint32); s2 = np.zeros((m, 92), np.uint32)
P = lambda x: x.ctypes.data
for rep in range(2):
    t0 = time.perf_counter()
    eng._ck(eng.lib.tecdsa_alice_proof_generate_batch(eng._ctx, ks.handle, P(ek_row), P(st_row), P(a_), P(c), P(r_), P(al), P(be), P(ga), P(ro),
                                                      P(z), P(e), P(s), P(s1), P(s2), m, 0), "gen")
    dt_g = time.perf_counter() - t0
    st = np.full(m, 255, np.uint8)
    t0 = time.perf_counter()
    eng._ck(eng.lib.tecdsa_alice_proof_verify_batch(eng._ctx, ks.handle, P(ek_row), P(st_row), P(c), P(z), P(e), P(s), P(s1), P(s2), P(st), m, 0), "ver")
    dt_v = time.perf_counter() - t0
ok = not st.any()
# 1 % tampered subset must reject
bad = rng.choice(m, size=m // 100, replace=False)
c2 = c.copy(); c2[bad, 5] ^= 1
st2 = np.full(m, 255, np.uint8)
eng._ck(eng.lib.tecdsa_alice_proof_verify_batch(eng._ctx, ks.handle, P(ek_row), P(st_row), P(c2), P(z), P(e), P(s), P(s1), P(s2), P(st2), m, 0), "ver")
ok = ok and bool((st2[bad] != 0).all()) and int((st2 != 0).sum()) == len(bad)
# byte parity of a sample against the oracle
allk = keyset + fixtures.load_keyset(1)
I = lambda row: int.from_bytes(row.tobytes(), "little")
for i in range(0, m, m // 16):
    ek = o.EncryptionKey(ns[ek_row[i]], ns[ek_row[i]] ** 2); lk = allk[st_row[i]]; stt = lk.h1_h2_n_tilde_vec[st_row[i] % 3]
    w = o.alice_proof_generate(I(a_[i]), I(c[i]), ek, stt, I(r_[i]), I(al[i]), I(be[i]), I(ga[i]), I(ro[i]))
    ok = ok and (I(z[i]), I(e[i]), I(s[i]), I(s1[i]), I(s2[i])) == (w.z, w.e, w.s, w.s1, w.s2)
W_PAIR = 2.73e8
print(json.dumps({"config": "configs[3]: 16k MtA range-proof (AliceProof) generate+verify", "count": m, "generate_s": dt_g, "verify_s": dt_v,
                  "pairs_per_s": m / (dt_g + dt_v), "achieved_tmac32_reference_oplist": W_PAIR * m / (dt_g + dt_v) / 1e12,
                  "tampered_rejected": len(bad), "sample_byte_parity_vs_oracle": 16, "ok": bool(ok)}), flush=True)
