"""Generates tests/golden/vectors_keygen.json — known answers of oracle/keygen_oracle.py
from fixed seeds over the committed key fixtures.  Same purpose as make_vectors.py: the restatement cannot drift silently,
the CUDA entry points (tests/test_keygen_gpu.py) get frozen bytes to be checked against, and a future build of the real
reference can be run on these inputs to pin the [R] encodings (zk-paillier hashing of N, salt, index; mask generation).
    python -m tests.golden.make_keygen_vectors
"""
import json
import os
import random

from oracle import gg20_oracle as o
from oracle import keygen_oracle as kg
from tests.golden import fixtures

HERE = os.path.dirname(os.path.abspath(__file__))
H = hex


def _prime(rng, b):
    while True:
        c = rng.getrandbits(b) | 1 | (1 << (b - 1))
        if all(c % s for s in (3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)) and all(pow(a, c - 1, c) == 1 for a in (2, 3, 5, 7)):
            return c


def build():
    keyset = fixtures.load_keyset(0)
    rng = random.Random(0xB20000F1)
    v = {"about": "oracle-generated known answers of the key-generation verification path, see make_keygen_vectors.py",
         "correct_key": [], "composite_dlog": [], "vss": []}
    for row in range(2):
        dk = keyset[row].dk
        n = dk.p * dk.q
        v["correct_key"].append({"row": row, "salt": kg.SALT_STRING.hex(), "rho": [H(x) for x in kg._rho_vec(n, kg.SALT_STRING)],
                                 "sigma": [H(x) for x in kg.correct_key_proof(dk)]})
    for bits in (512, 1024):
        p_t, q_t = _prime(rng, bits), _prime(rng, bits)
        phi = (p_t - 1) * (q_t - 1)
        h1 = rng.randrange(2, p_t * q_t)
        while True:
            xhi = rng.randrange(2, phi)
            try:
                pow(xhi, -1, phi)
                break
            except ValueError:
                continue
        nt, h1, h2, xn, xin = kg.h1_h2_n_tilde(p_t, q_t, h1, xhi)
        r1, r2 = rng.getrandbits(512), rng.getrandbits(512)
        pf1 = kg.composite_dlog_prove(o.DLogStatement(nt, h1, h2), xn, r1)
        pf2 = kg.composite_dlog_prove(o.DLogStatement(nt, h2, h1), xin, r2)
        v["composite_dlog"].append({"n_tilde": H(nt), "h1": H(h1), "h2": H(h2), "xhi_neg": H(xn), "xhi_inv_neg": H(xin), "r1": H(r1), "r2": H(r2),
                                    "proof_h1": [H(pf1.x), H(pf1.y)], "proof_h2": [H(pf2.x), H(pf2.y)]})
    for t, n in ((1, 3), (2, 5)):
        secret, coeff = rng.randrange(1, o.Q), [rng.randrange(1, o.Q) for _ in range(t)]
        vss, shares = kg.vss_share(t, n, secret, coeff)
        v["vss"].append({"t": t, "n": n, "secret": H(secret), "coefficients": [H(c) for c in coeff], "shares": [H(s) for s in shares],
                         "commitments": [[H(p[0]), H(p[1])] for p in vss.commitments]})
    # moduli with ONE small prime factor f and gcd(N, phi(N)) = 1: every sigma^N == rho check passes, so only the primorial
    # test gcd(P, N) == 1 (P = product of the primes <= 6379) can reject them; 6389 is the first prime P does not contain
    while True:
        big = _prime(rng, 2034)
        if all(big % f != 1 and (f - 1) % big != 0 for f in (3, 11, 6361, 6379, 6389)):
            break
    v["small_factor"] = {"q": H(big), "cases": []}
    for f, accept in ((3, False), (11, False), (6361, False), (6379, False), (6389, True)):
        dk = o.DecryptionKey(f, big)
        sigma = kg.correct_key_proof(dk)
        n = f * big
        assert kg.correct_key_verify(sigma, o.EncryptionKey(n, n * n)) == accept
        v["small_factor"]["cases"].append({"p": f, "accept": accept, "sigma": [H(x) for x in sigma]})
    return v


if __name__ == "__main__":
    with open(os.path.join(HERE, "vectors_keygen.json"), "w") as f:
        json.dump(build(), f, indent=1)
    print("wrote vectors_keygen.json")
