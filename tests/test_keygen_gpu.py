"""GPU parity tests of the key-generation verification entry points against
oracle/keygen_oracle.py.  EXPERIMENTAL: the CUDA side was written after the round's GPU budget was spent and has not run
on a GPU yet, so the module is skipped unless TECDSA_EXPERIMENTAL=1 (first thing to validate next round)."""
import os
import random

import numpy as np
import pytest

from oracle import gg20_oracle as o
from oracle import keygen_oracle as kg

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TECDSA_EXPERIMENTAL") != "1", reason="keygen verification kernels not yet validated on a GPU")]

OK, PROOF = 0, 10


def test_correct_key_verify_batch(engine, pkg, keyset):
    from mpecdsa_b200 import keygen
    ns, sig = [], []
    for k in keyset:
        ns.append(k.dk.p * k.dk.q)
        sig.append(kg.correct_key_proof(k.dk))
    # accepted proofs, then one tampered sigma, a proof against the wrong key, a wrong salt (separate call)
    bad = [list(s) for s in sig]
    bad[0][3] = (bad[0][3] + 1) % ns[0]
    bad[1] = sig[2]
    st = keygen.correct_key_verify(engine, ns + ns, sig + bad)
    want = [kg.correct_key_verify(s, o.EncryptionKey(n, n * n)) for n, s in zip(ns + ns, sig + bad)]
    assert want == [True, True, True, False, False, True]
    assert list(st) == [OK if w else PROOF for w in want]
    st = keygen.correct_key_verify(engine, ns, sig, salt=b"other")
    assert list(st) == [PROOF] * 3
    # a modulus sharing a factor with 6370 is rejected by the gcd test
    st = keygen.correct_key_verify(engine, [13 * ((ns[0] >> 4) | 1)], [sig[0]])
    assert list(st) == [PROOF]


def test_composite_dlog_verify_batch(engine, pkg, keyset):
    from mpecdsa_b200 import keygen
    from tests.test_keygen_oracle import _setup
    rng = random.Random(0xB2F3)
    stmts, proofs, want = [], [], []
    for i in range(6):
        p_t, q_t, h1, xhi = _setup(rng, bits=1024 if i < 2 else 512)
        nt, h1, h2, xn, xin = kg.h1_h2_n_tilde(p_t, q_t, h1, xhi)
        st1, st2 = o.DLogStatement(nt, h1, h2), o.DLogStatement(nt, h2, h1)
        pf1 = kg.composite_dlog_prove(st1, xn, rng.getrandbits(512))
        pf2 = kg.composite_dlog_prove(st2, xin, rng.getrandbits(512))
        stmts += [(nt, h1, h2), (nt, h2, h1), (nt, h1, h2), (nt, h1, h2)]
        proofs += [(pf1.x, pf1.y), (pf2.x, pf2.y), (pf1.x, pf1.y + 1), (pf2.x, pf2.y)]
        want += [True, True, False, kg.composite_dlog_verify(pf2, st1)]
    stmts.append((stmts[0][0], stmts[0][1] * 3, stmts[0][2])); proofs.append(proofs[0]); want.append(False)
    st = keygen.composite_dlog_verify(engine, stmts, proofs)
    assert [kg.composite_dlog_verify(kg.CompositeDLogProof(*p), o.DLogStatement(*s)) for s, p in zip(stmts, proofs)] == want
    assert list(st) == [OK if w else PROOF for w in want]


def test_vss_validate_share_batch(engine, pkg, keyset):
    from mpecdsa_b200 import keygen
    rng = random.Random(0xB2F4)
    comms, shares, idx, want = [], [], [], []
    for _ in range(3):                                 # one batch = one polynomial degree (t = 1, as in the fixtures)
        vss, sh = kg.vss_share(1, 5, rng.randrange(1, o.Q), [rng.randrange(1, o.Q)])
        for i, s in enumerate(sh, start=1):
            comms.append(vss.commitments); shares.append(s); idx.append(i); want.append(True)
        comms.append(vss.commitments); shares.append(sh[0]); idx.append(2); want.append(False)
        comms.append(vss.commitments); shares.append((sh[0] + 1) % o.Q); idx.append(1); want.append(False)
    st = keygen.vss_validate_share(engine, comms, shares, idx)
    assert list(st) == [OK if w else PROOF for w in want]
    vss, sh = kg.vss_share(3, 4, 12345, [7, 8, 9])
    st = keygen.vss_validate_share(engine, [vss.commitments] * 4, sh, [1, 2, 3, 4])
    assert list(st) == [OK] * 4
