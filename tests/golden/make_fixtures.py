"""Generates tests/golden/keys_t1n3.json — synthetic GG20 (t=1, n=3) `LocalKey` fixtures.

The reference has no fixture files; its tests build fresh keys per test with
`generate_init()` (/root/reference/src/utilities/mta/range_proofs.rs:592-613) and the keygen
protocol.  This script follows those recipes once, deterministically enough to be committed:
  - Paillier key: two 1024-bit primes p, q (Paillier::keypair, gg_2020/party_i.rs:162)
  - (N_tilde, h1, h2): N_tilde = p~ q~, h1 random, h2 = h1^xhi mod N_tilde with xhi invertible
    mod phi (range_proofs.rs:593-604)
  - Feldman shares of a degree-1 polynomial: x_i = f(i), X_i = x_i G, y = f(0) G
Primes come from the `cryptography` RSA key generator (OpenSSL).  Run:
    python -m tests.golden.make_fixtures
"""
import json
import os
import random

from cryptography.hazmat.primitives.asymmetric import rsa

from oracle import gg20_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))


def two_primes():
    k = rsa.generate_private_key(public_exponent=65537, key_size=2048)
    n = k.private_numbers()
    return n.p, n.q


def make_keyset(seed: int, n_parties: int = 3):
    rng = random.Random(seed)
    parties = []
    a0, a1 = rng.randrange(1, o.Q), rng.randrange(1, o.Q)       # f(x) = a0 + a1 x
    for i in range(1, n_parties + 1):
        p, q = two_primes()
        pt, qt = two_primes()
        nt = pt * qt
        phi = (pt - 1) * (qt - 1)
        h1 = rng.randrange(2, nt)
        while True:
            xhi = rng.randrange(2, phi)
            try:
                pow(xhi, -1, phi)
                break
            except ValueError:
                continue
        h2 = pow(h1, xhi, nt)
        x_i = (a0 + a1 * i) % o.Q
        parties.append({"i": i, "p": hex(p), "q": hex(q), "n_tilde": hex(nt), "h1": hex(h1), "h2": hex(h2),
                        "x_i": hex(x_i)})
    return {"t": 1, "n": n_parties, "secret": hex(a0), "parties": parties}


N_KEYSETS = 8          # "8 synthetic (t=1,n=3) key sets"


if __name__ == "__main__":
    # the primes come from OpenSSL's RNG, so a regeneration would change every key: key sets already in the file are KEPT
    # (the golden vectors of tests/golden/vectors_r01.json were produced with key sets 0 and 1) and only missing ones are added
    path = os.path.join(HERE, "keys_t1n3.json")
    sets = json.load(open(path))["keysets"] if os.path.exists(path) else []
    for s in range(len(sets), N_KEYSETS):
        sets.append(make_keyset(0xB2000005 + s))
    with open(path, "w") as f:
        json.dump({"about": "synthetic GG20 t=1,n=3 key sets; see make_fixtures.py", "keysets": sets}, f, indent=1)
    print("wrote", len(sets), "key sets")
