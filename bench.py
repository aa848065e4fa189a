#!/usr/bin/env python
"""bench.py — headline benchmark of the GG20 offline-signing arithmetic engine.

A "step" is one pass of the hot path over one batch of synthetic input.  At N=1 the
workload is BASELINE.json configs[1]: a batch of 65 536 independent 2048-bit Montgomery
modular exponentiations (2048-bit odd modulus, 2048-bit exponent), bit-exact vs GMP.
With N>1 every rank runs the same batch size on its own GPU (independent operands shard
with no data-path collective; "weak" scaling) and the ranks all-gather a fixed-size result
record (status summary + digest) over NCCL at the end of every step.

  value  : modexp/s, whole job, operands resident in HBM, CUDA-event timed, max over ranks.
  e2e    : same metric through the C ABI with HOST (pinned) buffers: H2D + kernel + D2H inside
           the timed region.
  roofline: the path is bound by the INT32 multiply-add pipe (IMAD.WIDE.U32), not HBM and not
           the tensor cores — `peak` is the on-box IMAD saturation
           micro-benchmark of the library (no integer entry exists in MEASURED_PEAKS.json);
           the HBM view (algorithmic bytes / time vs measured copy bandwidth) is reported
           beside it under "hbm".
  cpu_baseline: the oracle's C twin (GMP mpz_powm, the reference's default BigInt backend) on
           the host cores — a reported baseline, not the target.

`--impl reference` times that same CPU path as the reference arm (the Rust reference cannot be
built in this image: no cargo/rustc, crates not vendored).
"""
from __future__ import annotations

import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 65536
MOD_BITS = 2048
EXP_BITS = 2048
K = MOD_BITS // 32
EL = EXP_BITS // 32
SEED = 0xB2000002
# W_modexp(b,e) = 1.2 * e * (2k^2 + k) MAC32, k = b/32
W_MODEXP = 1.2 * EXP_BITS * (2 * K * K + K)
BYTES_PER_MODEXP = 4 * K * 4            # base, exponent, modulus in, result out (1 KiB)
WORKLOAD = "batch 64k 2048-bit Montgomery modexp on 1xB200, bit-exact vs GMP (BASELINE.json configs[1])"


def make_inputs(count: int, seed: int):
    """Random odd 2048-bit moduli with the top bit set, bases below 2^2048, 2048-bit exponents
    with the top bit set (fully-distinct-moduli variant)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 2**32, size=(count, K), dtype=np.uint32)
    exp = rng.integers(0, 2**32, size=(count, EL), dtype=np.uint32)
    mod = rng.integers(0, 2**32, size=(count, K), dtype=np.uint32)
    mod[:, 0] |= 1
    mod[:, K - 1] |= 0x80000000
    exp[:, EL - 1] |= 0x80000000
    return base, exp, mod


def load_oracle_lib():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libgg20_ref.so"))
    lib.oracle_gmp_version.restype = ctypes.c_char_p
    lib.oracle_modexp_batch.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return lib


def cpu_modexp(lib, base, exp, mod, threads: int):
    out = np.zeros_like(base)
    t0 = time.perf_counter()
    lib.oracle_modexp_batch(base.ctypes.data, exp.ctypes.data, mod.ctypes.data, None, out.ctypes.data,
                            base.shape[0], K, EL, threads)
    return out, time.perf_counter() - t0


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        loaded = sorted(x for x in sm if x > 0.5 * mx) or sorted(sm)
        return {"sm_mhz": loaded[len(loaded) // 2] if loaded else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}



# ---------------------------------------------------------------------------------------------
# Secondary workload: the full GG20 offline-signing stage (BASELINE.json metric part 1,
# configs[4] sharded per GPU).  SURVEY.md section 8(d): W_unit = 2.003e9 MAC32 (reference
# operation list), ~18 KB moved per unit.
OFFLINE_SESSIONS = 8192
W_UNIT = 2.003e9
W_UNIT_EXECUTED = 0.48e9      # MAC32 actually executed per unit after the declared shortcuts and the N-adic / p-adic forms
BYTES_PER_UNIT = 18 * 1024


def bench_offline(args, eng, pkg, torch, dist, rank, world, n_sessions):
    from mpecdsa_b200 import gg20
    from tests.golden import fixtures
    keysets = [fixtures.load_keyset(0), fixtures.load_keyset(1)]
    ks = gg20.KeySets(eng, keysets)
    sess, rnd = gg20.synthetic_batch(keysets, n_sessions, SEED + 17 * rank)
    U = 2 * n_sessions
    h_rnd = torch.from_numpy(rnd.view(np.int32)).pin_memory()
    h_sess = torch.from_numpy(sess.view(np.int32)).pin_memory()
    d_rnd, d_sess = h_rnd.cuda(), h_sess.cuda()
    d_status = torch.empty(U, dtype=torch.uint8, device="cuda")
    d_digest = torch.empty((U, 8), dtype=torch.int32, device="cuda")
    d_R = torch.empty((U, 16), dtype=torch.int32, device="cuda")
    d_sigma = torch.empty((U, 8), dtype=torch.int32, device="cuda")
    h_status = torch.empty(U, dtype=torch.uint8).pin_memory()
    h_digest = torch.empty((U, 8), dtype=torch.int32).pin_memory()
    h_R = torch.empty((U, 16), dtype=torch.int32).pin_memory()
    h_sigma = torch.empty((U, 8), dtype=torch.int32).pin_memory()
    record = torch.zeros((U, 9), dtype=torch.int32, device="cuda")       # status + digest per unit
    from mpecdsa_b200 import sharding

    def call(sessions, r, status, R, sigma, digest, mem):
        eng._ck(eng.lib.tecdsa_gg20_offline_batch(eng._ctx, ks.handle, sessions.data_ptr(), n_sessions, r.data_ptr(),
                                                  status.data_ptr(), R.data_ptr(), sigma.data_ptr(), None, digest.data_ptr(), mem),
                "gg20_offline_batch")

    gathered = [None]

    def step_device():
        call(d_sess, d_rnd, d_status, d_R, d_sigma, d_digest, pkg.DEVICE)
        record[:, 0] = d_status
        record[:, 1:] = d_digest
        gathered[0] = sharding.gather_records(record, world)              # the single NCCL all-gather

    def step_host():
        call(h_sess, h_rnd, h_status, h_R, h_sigma, h_digest, pkg.HOST)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    steps = max(1, min(args.steps, 3))
    step_device(); barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launch_count()
    ev0.record()
    for _ in range(steps):
        step_device()
    ev1.record(); barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    k_ms, k_launches = eng.last_kernel_ms()
    step_host(); barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_host()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([dev_ms, e2e_ms, k_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, k_ms = (float(x) for x in t.tolist())
    all_ok = bool((gathered[0][:, :, 0] == 0).all().item()) and bool((h_status == 0).all().item())
    same = bool(torch.equal(h_digest, d_digest.cpu()))
    ks.free()
    total_units = U * world
    return {
        "metric": "GG20 (t=1,n=3) offline-signing phases/s", "value": total_units * steps / (dev_ms * 1e-3), "unit": "phases/s",
        "units_per_gpu": U, "sessions_per_gpu": n_sessions, "steps": steps, "ms_per_step": dev_ms / steps,
        "e2e": {"value": total_units * steps / (e2e_ms * 1e-3), "unit": "phases/s", "h2d_bytes_per_step": int(rnd.nbytes + sess.nbytes),
                "d2h_bytes_per_step": int(U * (1 + 32 + 64 + 32)), "ms_per_step": e2e_ms / steps},
        "gpu_launches_per_step": int(launches // steps), "kernels_ms_per_step": k_ms,
        "all_units_ok": all_ok, "host_and_device_paths_agree": same,
        "work": {"W_unit_reference_oplist_mac32": W_UNIT, "W_unit_executed_mac32": W_UNIT_EXECUTED,
                 "achieved_reference_oplist_tmac32": W_UNIT * U / (k_ms * 1e-3) / 1e12,
                 "achieved_executed_tmac32": W_UNIT_EXECUTED * U / (k_ms * 1e-3) / 1e12,
                 "hbm_algorithmic_gbs": BYTES_PER_UNIT * U / (k_ms * 1e-3) / 1e9},
    }


def cpu_unit_baseline(threads: int):
    lib = load_oracle_lib()
    lib.oracle_unit_oplist.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64]
    units = max(threads, 8)
    t0 = time.perf_counter()
    lib.oracle_unit_oplist(units, threads, 7)
    dt = time.perf_counter() - t0
    return {"value": units / dt, "unit": "phases/s", "cores": threads, "kind": "port",
            "sample": f"{units} units: the reference's big-integer operation list of one OfflineStage (76 mpz_powm + 22 mpz_invert at the "
                      f"reference's operand sizes, redundant verifications included; EC/hash <1 % omitted), GMP {lib.oracle_gmp_version().decode()}"}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (GMP mpz_powm behind
    BigInt::mod_pow) on all host threads; each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib = load_oracle_lib()
    threads = host_threads()
    sample = max(threads * 16, 512)
    base, exp, mod = make_inputs(sample, SEED)
    for _ in range(args.warmup):
        cpu_modexp(lib, base[: threads * 2], exp[: threads * 2], mod[: threads * 2], threads)
    t = 0.0
    for _ in range(args.steps):
        _, dt = cpu_modexp(lib, base, exp, mod, threads)
        t += dt
    value = sample * args.steps / t
    line = {
        "impl": "reference", "metric": "2048-bit modexp/s", "value": value, "unit": "modexp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "mod_bits": MOD_BITS, "exp_bits": EXP_BITS,
                   "note": "CPU arm: oracle C twin over GMP %s (the backend of the reference's default feature); the Rust "
                           "reference itself cannot be built here (no cargo/rustc, crates not vendored)" % lib.oracle_gmp_version().decode()},
        "cpu_baseline": {"value": value, "unit": "modexp/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} modexps per step x {args.steps} steps, cpu={cpu_model()}"},
        "e2e": {"value": value, "unit": "modexp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--offline-sessions", type=int, default=OFFLINE_SESSIONS, help="sessions per GPU for the offline-stage section (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = entry.load_package()
    stream = torch.cuda.current_stream()
    eng = pkg.Engine(local_rank, stream.cuda_stream)
    batch = args.batch

    # ---- inputs: rank-distinct synthetic operands, one device-resident copy, one pinned host copy
    base, exp, mod = make_inputs(batch, SEED + rank)
    h_base, h_exp, h_mod = (torch.from_numpy(x.view(np.int32)).pin_memory() for x in (base, exp, mod))
    h_out = torch.empty((batch, K), dtype=torch.int32).pin_memory()
    h_st = torch.empty(batch, dtype=torch.uint8).pin_memory()
    d_base, d_exp, d_mod = (x.cuda(non_blocking=True) for x in (h_base, h_exp, h_mod))
    d_out = torch.empty((batch, K), dtype=torch.int32, device="cuda")
    d_st = torch.empty(batch, dtype=torch.uint8, device="cuda")
    record = torch.zeros(8, dtype=torch.int64, device="cuda")            # per-rank result record
    gathered = torch.zeros(8 * world, dtype=torch.int64, device="cuda")

    def step_device():
        eng.modexp_raw(MOD_BITS, EL, d_base, d_exp, d_mod, d_out, d_st, mem=pkg.DEVICE)
        if world > 1:
            record[0] = d_st.sum()
            record[1:5] = d_out[:, :4].to(torch.int64).sum(dim=0)
            dist.all_gather_into_tensor(gathered, record)

    def step_host():
        eng.modexp_raw(MOD_BITS, EL, h_base, h_exp, h_mod, h_out, h_st, mem=pkg.HOST)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing (value) + per-launch kernel time (roofline)
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
        kernel_ms.append(None)
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    # kernel-only duration of one launch (events recorded by the library around the kernel)
    step_device(); torch.cuda.synchronize()
    k_ms, k_launches = eng.last_kernel_ms()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end timing through the C ABI with host buffers
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([dev_ms, e2e_s * 1e3, k_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, k_ms = (float(x) for x in t.tolist())

    # ---- parity spot check inside the bench (not timed): sample vs the GMP oracle
    ok = bool((d_st == 0).all().item()) and bool((h_st == 0).all().item())
    if rank == 0:
        lib = load_oracle_lib()
        idx = np.linspace(0, batch - 1, 64).astype(np.int64)
        want, _ = cpu_modexp(lib, np.ascontiguousarray(base[idx]), np.ascontiguousarray(exp[idx]),
                             np.ascontiguousarray(mod[idx]), 4)
        got_dev = d_out.cpu().numpy().view(np.uint32)[idx]
        got_host = h_out.numpy().view(np.uint32)[idx]
        ok = ok and np.array_equal(want, got_dev) and np.array_equal(want, got_host)

    offline = None
    if args.offline_sessions > 0:
        try:
            offline = bench_offline(args, eng, pkg, torch, dist, rank, world, args.offline_sessions)
        except Exception as exc:          # the secondary section must not take the headline line down with it
            offline = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total = batch * world
    value = total * args.steps / (dev_ms * 1e-3)
    e2e_value = total * args.steps / (e2e_ms * 1e-3)
    peak_mac, peak_ms = eng.imad_peak()
    achieved_mac = W_MODEXP * batch / (k_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
    hbm_achieved = BYTES_PER_MODEXP * batch / (k_ms * 1e-3) / 1e9

    cpu = None
    if not args.no_cpu_baseline:
        lib = load_oracle_lib()
        threads = host_threads()
        sample = max(threads * 48, 1024)                 # ~3 ms each -> 10-30 s of CPU work in total
        sb, se, sm_ = make_inputs(sample, SEED + 999)
        _, dt = cpu_modexp(lib, sb, se, sm_, threads)
        _, dt1 = cpu_modexp(lib, sb[:64], se[:64], sm_[:64], 1)
        cpu = {"value": sample / dt, "unit": "modexp/s", "cores": threads, "kind": "port",
               "sample": f"{sample} modexps of the same shape on {threads} threads (GMP {lib.oracle_gmp_version().decode()} mpz_powm, "
                         f"{cpu_model()}); single thread: {64 / dt1:.1f} modexp/s"}

    line = {
        "metric": "2048-bit modexp/s", "value": value, "unit": "modexp/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": batch, "mod_bits": MOD_BITS, "exp_bits": EXP_BITS,
                   "distinct_moduli": True, "window_bits": 5, "parallelism": f"shard{world}",
                   "cache": "per-step working set (48 MiB operands + 512 MiB window tables) exceeds the 126 MB L2",
                   "parity_ok": ok},
        "e2e": {"value": e2e_value, "unit": "modexp/s", "h2d_bytes_per_step": int(batch * (2 * K + EL) * 4),
                "d2h_bytes_per_step": int(batch * (K * 4 + 1)), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "int32-mad", "achieved": achieved_mac / 1e12, "peak": peak_mac / 1e12, "unit": "TMAC32/s",
                     "frac": achieved_mac / peak_mac,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the ncu --set full capture of this kernel
                     # (profiles/r01_ncu_modexp2048_summary.md): window tables written once and re-read; algorithmic bytes are 64 MiB
                     "traffic": 1.863e9 if batch == BATCH else None,
                     "kernel": "modexp_kernel<64,TPI>", "kernel_ms": k_ms, "launches_per_step": k_launches,
                     "work_per_launch_mac32": W_MODEXP * batch,
                     "peak_source": "on-box IMAD.WIDE.U32 saturation micro-benchmark (tecdsa_imad_peak); "
                                    "MEASURED_PEAKS.json has no integer entry",
                     "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                             "peak_source": hbm_src, "algorithmic_bytes_per_launch": BYTES_PER_MODEXP * batch}},
        "cpu_baseline": cpu,
        "clocks": clocks,
        "offline_stage": offline,
    }
    if offline is not None and "error" not in offline:
        offline["roofline_frac_reference_oplist"] = offline["work"]["achieved_reference_oplist_tmac32"] * 1e12 / peak_mac
        offline["roofline_frac_executed"] = offline["work"]["achieved_executed_tmac32"] * 1e12 / peak_mac
        if not args.no_cpu_baseline:
            offline["cpu_baseline"] = cpu_unit_baseline(host_threads())
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("parity check against the oracle FAILED")
    if offline is not None and "error" in offline:
        raise SystemExit("offline stage section failed: " + offline["error"])
    if offline is not None and not (offline["all_units_ok"] and offline["host_and_device_paths_agree"]):
        raise SystemExit("offline stage: a unit failed or the host/device paths disagree")


if __name__ == "__main__":
    main()
