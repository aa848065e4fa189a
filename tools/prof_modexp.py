"""Tiny driver for ncu: N launches of the config-2 modexp batch through the C ABI (device buffers)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as entry
import bench
pkg = entry.load_package()
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
tpi = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
K = bits // 32
rng = np.random.default_rng(1)
base = rng.integers(0, 2**32, size=(count, K), dtype=np.uint32)
exp = rng.integers(0, 2**32, size=(count, 64), dtype=np.uint32)
mod = rng.integers(0, 2**32, size=(count, K), dtype=np.uint32); mod[:, 0] |= 1; mod[:, K-1] |= 0x80000000
eng = pkg.Engine(0)
eng.set_tpi(bits, tpi)
t = lambda a: torch.from_numpy(a.view(np.int32)).cuda()
B, E, M = t(base), t(exp), t(mod)
out = torch.zeros_like(B); st = torch.zeros(count, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    eng.modexp_raw(bits, 64, B, E, M, out, st, mem=pkg.DEVICE)
    eng.sync()
    print("kernel ms", eng.last_kernel_ms())
