"""First-contact GPU check: modexp correctness for every (mod_bits, TPI) shape + timing + IMAD peak."""
import ctypes, os, random, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "..", "multi-party-ecdsa_b200", "libtecdsa_b200.so"))
lib.tecdsa_last_error.restype = ctypes.c_char_p
ctx = ctypes.c_void_p()
rc = lib.tecdsa_ctx_create(ctypes.byref(ctx), 0, None)
assert rc == 0, lib.tecdsa_last_error()

def to_limbs(vals, K):
    b = b"".join(int(v).to_bytes(K * 4, "little") for v in vals)
    return np.frombuffer(b, dtype=np.uint32).reshape(len(vals), K).copy()
def from_limbs(a):
    return [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]
def ptr(a): return a.ctypes.data_as(ctypes.c_void_p)

def run(bits, tpi, count, ebits, check=True, rng=random.Random(1)):
    K = bits // 32; EL = (ebits + 31) // 32
    mods = [rng.getrandbits(bits) | (1 << (bits - 1)) | 1 for _ in range(count)]
    bases = [rng.getrandbits(bits) % m for m in mods]
    exps = [rng.getrandbits(ebits) for _ in range(count)]
    if count > 8:
        exps[0] = 0; exps[1] = 1; bases[2] = 0; bases[3] = mods[3] - 1; mods[4] = (1 << bits) - 1; bases[4] %= mods[4]
        mods[5] = (1 << (bits - 1)) + 1; bases[5] %= mods[5]; mods[6] = 3; bases[6] = 2; bases[7] = (1<<bits)-1
    B, E, M = to_limbs(bases, K), to_limbs(exps, EL), to_limbs(mods, K)
    out = np.zeros_like(B); st = np.full(count, 255, dtype=np.uint8)
    assert lib.tecdsa_ctx_set_tpi(ctx, bits, tpi) == 0
    t0 = time.time()
    rc = lib.tecdsa_modexp_batch(ctx, bits, EL, ptr(B), ptr(E), ptr(M), None, ctypes.c_size_t(0), ptr(out), ptr(st), ctypes.c_size_t(count), 0)
    dt = time.time() - t0
    assert rc == 0, lib.tecdsa_last_error()
    ms = ctypes.c_float(); nl = ctypes.c_int()
    lib.tecdsa_ctx_last_kernel_ms(ctx, ctypes.byref(ms), ctypes.byref(nl))
    bad = 0
    if check:
        got = from_limbs(out)
        for i in range(count):
            if got[i] != pow(bases[i], exps[i], mods[i]) or st[i] != 0:
                bad += 1
                if bad < 4: print("  MISMATCH", bits, tpi, i, hex(got[i])[:40], hex(pow(bases[i], exps[i], mods[i]))[:40], st[i])
    print(f"bits={bits} tpi={tpi} count={count} ebits={ebits} kernel_ms={ms.value:.3f} wall_s={dt:.3f} rate={count/ms.value*1e3:.0f}/s bad={bad}", flush=True)
    return bad

peak = ctypes.c_double(); pms = ctypes.c_float()
assert lib.tecdsa_imad_peak(ctx, ctypes.byref(peak), ctypes.byref(pms)) == 0, lib.tecdsa_last_error()
print(f"imad_peak {peak.value/1e12:.3f} TMAC32/s ({pms.value:.3f} ms)", flush=True)
bad = 0
for bits, tpis in ((2048, (4, 8, 16, 32)), (1024, (4, 8, 16)), (4096, (8, 16, 32))):
    for tpi in tpis:
        bad += run(bits, tpi, 300, 200)
        bad += run(bits, tpi, 37, bits)
print("correctness bad =", bad, flush=True)
if "--time" in sys.argv:
    for tpi in (4, 8, 16, 32):
        run(2048, tpi, 65536, 2048, check=False)
    for tpi in (8, 16, 32):
        run(4096, tpi, 32768, 2048, check=False)
    for tpi in (4, 8, 16):
        run(1024, tpi, 65536, 1024, check=False)
sys.exit(1 if bad else 0)
