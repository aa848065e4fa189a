"""Device-resident throughput of the secp256k1 scalar multiplications (BASELINE configs[2] shape: 2^20 scalars): generator base through
the 8-bit fixed-base tables and variable base through GLV + signed 5-bit windows, timed from the host around call + stream sync (device buffers, one launch of >= 15 ms: the ~20 us of launch and sync latency are < 0.2 %).
    python tools/ec_throughput.py [log2_count]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as entry
pkg = entry.load_package()
pkg._bind_l01(pkg.load_library())
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
rng = np.random.default_rng(0xB2000003)
k = rng.integers(0, 2**32, size=(n, 8), dtype=np.uint32); k[:, 7] &= 0x7FFFFFFF; k[:, 0] |= 1
eng = pkg.Engine(0)
t = lambda a: torch.from_numpy(a.view(np.int32)).cuda()
K, K2 = t(k), t(np.roll(k, 1, axis=0))
out = torch.zeros((n, 16), dtype=torch.int32, device="cuda"); out2 = torch.zeros_like(out)
res = {"count": n}
for name, pts, sc, dst in (("generator_base", None, K, out), ("variable_base", out, K2, out2)):
    ms = []
    for _ in range(3):
        eng.sync(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng._ck(eng.lib.tecdsa_secp_mul_batch(eng._ctx, None if pts is None else pts.data_ptr(), sc.data_ptr(), dst.data_ptr(), n, pkg.DEVICE), "secp_mul")
        eng.sync()
        ms.append((time.perf_counter() - t0) * 1e3)
    res[name] = {"ms": ms, "mul_per_s": n / (min(ms) * 1e-3)}
print(json.dumps(res))
