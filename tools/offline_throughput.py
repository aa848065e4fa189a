"""Throughput of the batched GG20 offline stage on one GPU (host buffers in, host results out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry
from tests.golden import fixtures
pkg = entry.load_package()
from mpecdsa_b200 import gg20
keysets = [fixtures.load_keyset(0), fixtures.load_keyset(1)]
eng = pkg.Engine(0)
ks = gg20.KeySets(eng, keysets)
for n_sessions in [int(x) for x in (sys.argv[1:] or ["1024", "4096"])]:
    sess, rnd = gg20.synthetic_batch(keysets, n_sessions, 1234)
    for rep in range(2):
        t0 = time.perf_counter()
        res = gg20.offline_batch(eng, ks, sess, rnd)
        dt = time.perf_counter() - t0
        ms, nl = eng.last_kernel_ms()
        bad = int((res.status != 0).sum())
        print(f"sessions={n_sessions} units={2*n_sessions} wall={dt:.3f}s kernels={ms:.1f}ms launches={nl} units/s={2*n_sessions/dt:.0f} (device {2*n_sessions/ms*1e3:.0f}) bad={bad}", flush=True)
