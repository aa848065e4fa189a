"""Deterministic sampling of every value the reference draws from the OS RNG on the hot path
— TEST INFRASTRUCTURE ONLY (see gg20_oracle.py).  The engine and the oracle both take the
sampled values as explicit inputs, so only the RANGES matter for parity, not the generator.
Ranges follow the reference line by line."""
import hashlib
from math import gcd

from . import gg20_oracle as o


class Drbg:
    """SHA-256 counter generator (not the reference's RNG; see module docstring)."""

    def __init__(self, seed: int, label: str = ""):
        self.key = hashlib.sha256(f"{seed:x}/{label}".encode()).digest()
        self.ctr = 0

    def bits(self, n: int) -> int:
        out = b""
        while len(out) * 8 < n:
            out += hashlib.sha256(self.key + self.ctr.to_bytes(8, "big")).digest()
            self.ctr += 1
        return int.from_bytes(out, "big") >> (len(out) * 8 - n)

    def below(self, bound: int) -> int:
        """`BigInt::sample_below`: rejection sampling on bit_length(bound) bits."""
        n = bound.bit_length()
        while True:
            v = self.bits(n)
            if v < bound:
                return v

    def scalar(self) -> int:
        """`Scalar::<Secp256k1>::random()` (non-zero)."""
        while True:
            v = self.below(o.Q)
            if v:
                return v

    def unit_mod(self, n: int) -> int:
        """`BigInt::from_paillier_key` / from_modulo (range_proofs.rs:543-556): r < N, gcd(r,N)=1."""
        while True:
            v = self.below(n)
            if gcd(v, n) == 1:
                return v


def sample_unit(rng: Drbg, keys, s_l, pos: int) -> o.UnitRandomness:
    """All randomness of signer position `pos` in one offline session (order of use)."""
    q3 = o.Q ** 3
    lk = keys[pos]
    l_s = [x - 1 for x in s_l]
    other = 1 - pos
    n_own = lk.paillier_key_vec[lk.i - 1].n
    n_peer = lk.paillier_key_vec[l_s[other]].n
    r = o.UnitRandomness()
    r.gamma_i = rng.scalar()                                   # party_i.rs:563
    r.k_i = rng.scalar()                                       # party_i.rs:565
    r.blind = rng.bits(256)                                    # party_i.rs:574 BigInt::sample(SECURITY)
    r.r_k = rng.below(n_own)                                   # mta/mod.rs:57
    for st in lk.h1_h2_n_tilde_vec:                            # range_proofs.rs:48-51
        r.alice.append((rng.below(q3), rng.unit_mod(n_own), rng.below(q3 * st.N), rng.below(o.Q * st.N)))
    r.beta_tag_gamma = rng.below(n_peer); r.r_gamma = rng.below(n_peer)      # mta/mod.rs:97-98
    r.nonce_gamma_b = rng.scalar(); r.nonce_gamma_beta = rng.scalar()
    r.beta_tag_w = rng.below(n_peer); r.r_w = rng.below(n_peer)
    r.nonce_w_b = rng.scalar(); r.nonce_w_beta = rng.scalar()
    r.l = rng.scalar()                                         # party_i.rs:628
    r.ped_s1 = rng.scalar(); r.ped_s2 = rng.scalar()
    st = lk.h1_h2_n_tilde_vec[l_s[other]]                      # zk_pdl_with_slack/mod.rs:73-77
    beta = 1 + rng.below(n_own - 2)                            # sample_range(1, n-1)
    r.pdl = (rng.below(q3), beta, rng.below(o.Q * st.N), rng.below(q3 * st.N))
    r.heg_s1 = rng.scalar(); r.heg_s2 = rng.scalar()
    return r
