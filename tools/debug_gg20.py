"""Field-by-field comparison of one GPU offline session against the oracle (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry
from oracle import gg20_oracle as o
from oracle.sampling import Drbg, sample_unit
from tests.golden import fixtures

pkg = entry.load_package()
from mpecdsa_b200 import gg20

keyset = fixtures.load_keyset()
a, b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 1)
s_l = [a + 1, b + 1]
keys = [keyset[a], keyset[b]]
rng = Drbg(5, "dbg")
rnd = [sample_unit(rng, keys, s_l, p) for p in range(2)]
eng = pkg.Engine(0)
ks = gg20.KeySets(eng, [keyset])
res = gg20.offline_batch(eng, ks, [(0, a, b)], gg20.pack_randomness(rnd))
print("status", list(res.status))
U = 2
l_s = [a, b]
exp = [dict(), dict()]
Q = o.Q
pt = gg20.pack_point
w, gamma, k, gg, m_a = [], [], [], [], []
for p in range(2):
    lk, r = keys[p], rnd[p]
    li = o.lagrange_at_zero(l_s[p], l_s)
    w.append(li * lk.x_i % Q); gamma.append(r.gamma_i); k.append(r.k_i)
    gg.append(o.pt_mul(o.G, r.gamma_i))
    ek = lk.paillier_key_vec[lk.i - 1]
    ma = o.message_a(r.k_i, ek, r.r_k, lk.h1_h2_n_tilde_vec, r.alice)
    m_a.append(ma)
    e = exp[p]
    e["W"] = w[p]; e["GG"] = pt(gg[p]); e["COM"] = o.hash_commitment(o.bn_from_bytes(o.pt_compress(gg[p])), r.blind)
    e["MK"] = 1 + r.k_i * ek.n; e["CK"] = ma.c
    for x in range(3):
        al, be, ga, ro = r.alice[x]
        st = lk.h1_h2_n_tilde_vec[x]
        e[f"ALIN{x}"] = al * ek.n + 1
        e[f"U{x}"] = (al * ek.n + 1) * pow(be, ek.n, ek.nn) % ek.nn
        e[f"Z{x}"] = ma.range_proofs[x].z
        e[f"WP{x}"] = pow(st.g, al, st.N) * pow(st.ni, ga, st.N) % st.N
        e[f"E{x}"] = ma.range_proofs[x].e; e[f"S{x}"] = ma.range_proofs[x].s
        e[f"S1{x}"] = ma.range_proofs[x].s1; e[f"S2{x}"] = ma.range_proofs[x].s2
mb = [[None, None], [None, None]]
beta = [[0, 0], [0, 0]]
for p in range(2):
    q_ = 1 - p
    lk, r = keys[p], rnd[p]
    ek_o = lk.paillier_key_vec[l_s[q_]]
    e = exp[p]
    for x in range(3):
        pf = m_a[q_].range_proofs[x]; st = lk.h1_h2_n_tilde_vec[x]
        ze = pow(pf.z, pf.e, st.N); ce = pow(m_a[q_].c, pf.e, ek_o.nn)
        e[f"ZE{x}"] = ze; e[f"CE{x}"] = ce; e[f"ZEI{x}"] = pow(ze, -1, st.N); e[f"CEI{x}"] = pow(ce, -1, ek_o.nn)
        e[f"GS1{x}"] = (pf.s1 * ek_o.n + 1) % ek_o.nn
        e[f"WV{x}"] = pow(st.g, pf.s1, st.N) * pow(st.ni, pf.s2, st.N) * e[f"ZEI{x}"] % st.N
        e[f"UV{x}"] = e[f"GS1{x}"] * pow(pf.s, ek_o.n, ek_o.nn) * e[f"CEI{x}"] % ek_o.nn
    rb = o.message_b(gamma[p], ek_o, m_a[q_], r.r_gamma, r.beta_tag_gamma, lk.h1_h2_n_tilde_vec, r.nonce_gamma_b, r.nonce_gamma_beta)
    rw = o.message_b(w[p], ek_o, m_a[q_], r.r_w, r.beta_tag_w, lk.h1_h2_n_tilde_vec, r.nonce_w_b, r.nonce_w_beta)
    mb[p][0], beta[p][0] = rb; mb[p][1], beta[p][1] = rw
    e["LBG"] = 1 + r.beta_tag_gamma * ek_o.n; e["CBG"] = rb[0].c; e["CBW"] = rw[0].c
    e["BETA_G"] = rb[1]; e["NU"] = rw[1]
    for i, d in enumerate((rb[0].b_proof, rb[0].beta_tag_proof, rw[0].b_proof, rw[0].beta_tag_proof)):
        e[f"DL{i}"] = pt(d.pk) | (pt(d.pk_t_rand_commitment) << 512) | (d.challenge_response << 1024)
for p in range(2):
    q_ = 1 - p
    lk, r = keys[p], rnd[p]
    e = exp[p]
    for nm, msg in (("G", mb[q_][0]), ("W", mb[q_][1])):
        e[f"DP{nm}"] = pow(msg.c % lk.dk.p ** 2, lk.dk.p - 1, lk.dk.p ** 2)
        e[f"DQ{nm}"] = pow(msg.c % lk.dk.q ** 2, lk.dk.q - 1, lk.dk.q ** 2)
    al = o.verify_proofs_get_alpha(mb[q_][0], lk.dk, k[p]); mu = o.verify_proofs_get_alpha(mb[q_][1], lk.dk, k[p])
    e["ALPHA"] = al[0]; e["MU"] = mu[0]
    e["DELTA"] = (k[p] * gamma[p] + al[0] + beta[p][0]) % Q
    e["SIGMA"] = (k[p] * w[p] + mu[0] + beta[p][1]) % Q
    e["T"] = pt(o.pt_add(o.pt_mul(o.G, e["SIGMA"]), o.pt_mul(o.H2, r.l)))
want = o.offline_session(keys, s_l, rnd)
for p in range(2):
    exp[p]["R"] = pt(want[p].R); exp[p]["SIGMA"] = want[p].sigma_i
    exp[p]["DIGEST"] = int.from_bytes(want[p].transcript, "big")
order = ["W", "GG", "COM", "MK", "ALIN0", "CK", "U0", "U1", "U2", "Z0", "Z1", "Z2", "WP0", "WP1", "WP2", "E0", "E1", "E2",
         "S10", "S20", "S0", "S1", "S2", "GS10", "ZE0", "ZEI0", "CEI0", "ZEI1", "CEI1", "ZEI2", "CEI2", "WV0", "UV0", "WV1", "UV1", "WV2", "UV2",
         "LBG", "CBG", "CBW", "BETA_G", "NU", "DL0", "DL1", "DL2", "DL3", "DPG", "DQG", "DPW", "DQW", "ALPHA", "MU", "DELTA", "SIGMA", "T", "R", "DIGEST"]
bad = 0
for name in order:
    got = gg20.debug_field(eng, name, U)
    gi = pkg.limbs_to_ints(got)
    for p in range(2):
        if name in exp[p]:
            ok = gi[p] == exp[p][name]
            if not ok:
                bad += 1
                print(f"MISMATCH {name} unit {p}: got {hex(gi[p])[:50]}.. want {hex(exp[p][name])[:50]}..")
fl = gg20.debug_field(eng, "FLAGS", U)
print("flags", [list(fl[p].view(np.uint8)[:11]) for p in range(2)])
print("mismatches:", bad)
