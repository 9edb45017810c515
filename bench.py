#!/usr/bin/env python
"""bench.py — events/sec of the segmented event fold (BASELINE.json metric) on N B200s of one node.

A "step" is one full pass of the hot path over one batch of synthetic input: rebuilding every
aggregate's state from its CSR event log (configs[1]: 1,048,576 aggregates x 32 fixed 64-byte
events = 2 GiB of events per GPU; the log is far larger than the 126 MB L2, so no flush is needed
between timed iterations).

  value   whole-job events/s with the log resident in HBM, K pipelined folds, CUDA events on the
          engine's stream, max over ranks (weak scaling: every rank folds its own shard of
          aggregates, no data-path collective — aggregates are independent units)
  e2e     the same metric through the C ABI with HOST buffers: every step copies the log from pinned
          host memory (sgr_load_events), folds (sgr_fold) and reads the state table back
          (sgr_export_states)
  roofline   algorithmic bytes / device time of the fold kernel against the measured HBM peak
  cpu_baseline   the CPU oracle (port of the reference's fold) on this box's host cores

`--impl reference` times the reference's CPU implementation of the path instead (the oracle port:
the reference is Scala/JVM and cannot be built in this image), rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_AGG = 1 << 20
EVENTS_PER_AGG = 32
STATE_BYTES = 16
METRIC = "events/sec replayed (segmented per-aggregate event fold)"
WORKLOAD = "configs[1]: 1,048,576 aggregates x 32 fixed-width 64-B events, single B200 segmented fold (per GPU)"


def algorithmic_bytes(n_agg: int, epa: int) -> int:
    """B_alg = stored event bytes + 8*(nAgg+1) CSR offsets + S*nAgg states written."""
    return n_agg * epa * 64 + 8 * (n_agg + 1) + STATE_BYTES * n_agg


def measured_peak_gbs() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Polls NVML for SM clock and throttle reasons while the timed regions run."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.samples.append((mhz, util))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        mhz = [m for m, _ in self.samples]
        return {"sm_mhz": int(statistics.median(mhz)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(mhz)}


def host_config2_log(n_agg: int, epa: int, seed: int):
    """The config-2 Counter log built on the host (numpy), for the CPU legs."""
    import numpy as np

    from surge_b200 import formats as F

    rng = np.random.Generator(np.random.Philox(seed))
    n = n_agg * epa
    rec = np.zeros((n, 16), dtype=np.int32)
    u = rng.random(n, dtype=np.float32)
    rec[:, 0] = np.where(u < 0.45, F.COUNT_INCREMENTED, np.where(u < 0.9, F.COUNT_DECREMENTED, F.NO_OP_EVENT))
    idx = np.arange(n, dtype=np.int64)
    rec[:, 1] = (idx % epa + 1).astype(np.int32)
    rec[:, 2] = (idx // epa).astype(np.int32)
    rec[:, 4] = rng.integers(0, 1 << 31, size=n, dtype=np.int64).astype(np.int32)
    off = (np.arange(n_agg + 1, dtype=np.uint64) * np.uint64(epa * 64))
    return rec, off


def time_cpu_oracle(rec, off, threads: int, min_seconds: float, max_reps: int):
    from oracle import oracle as O

    O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec[: 64 * 16], off[:33], threads=1)  # load the library
    reps, t_total, nev = 0, 0.0, 0
    while reps < max_reps and (reps == 0 or t_total < min_seconds):
        t0 = time.perf_counter()
        _, n, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, threads=threads)
        t_total += time.perf_counter() - t0
        nev += n
        reps += 1
    return nev / t_total, reps, t_total


def run_reference(args) -> None:
    """The reference's CPU implementation of the path (oracle port, all host threads), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # one step = one pass over the full configs[1] log (33.5 M events, 2 GiB, far larger than any CPU cache, like the
    # GPU arm's step); a cache-resident sample would overstate what the CPU path does on this workload
    n_agg = N_AGG
    rec, off = host_config2_log(n_agg, EVENTS_PER_AGG, seed=2)
    from oracle import oracle as O

    for _ in range(max(args.warmup, 1)):
        O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, threads=cores)
    t0 = time.perf_counter()
    nev = 0
    for _ in range(args.steps):
        _, n, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, threads=cores)
        nev += n
    dt = time.perf_counter() - t0
    value = nev / dt
    sample = f"{n_agg} aggregates x {EVENTS_PER_AGG} events per step (the full configs[1] log, pageable host memory), {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "impl_note": "CPU port of the reference's fold (oracle/sgr_oracle.c); the Scala/JVM reference cannot be built in this image"},
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


ROUTED_AGG_PER_GPU = 1_250_000   # configs[2]: 10 M aggregates x 100 events over 8 GPUs = 1.25 M x 100 per GPU
ROUTED_EPA = 100


def routed_pipeline(rank, world, local_rank, dev, barrier, note, strong=False):
    """configs[2]-shaped per-GPU work (weak: 1.25 M aggregates x 100 events of 64 B originate on every rank):
    route (K4) + exchange (NCCL all-to-all, then fused peer-memory scatter) + stable group-by (K5) + fold.
    Stage times are CUDA-event times on the engine's stream, max over ranks; the job rate uses the wall time
    of the whole call, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from surge_b200 import ReplayEngine
    from surge_b200 import dist as D
    from surge_b200 import programs as P

    n_global = ROUTED_AGG_PER_GPU * (8 if strong else world)   # strong: always the full 10 M x 100 problem
    epa = ROUTED_EPA
    # this rank's source partitions hold the aggregates g with g % world == rank, in arrival order
    # (event k of every aggregate before event k+1: aggregates interleaved, per-aggregate order kept)
    g_mine = torch.arange(rank, n_global, world, device=dev, dtype=torch.int64)
    n = g_mine.numel() * epa
    gen = torch.Generator(device=dev)
    gen.manual_seed(1000 + rank)
    r = torch.zeros((n, 16), dtype=torch.int32, device=dev)
    na = g_mine.numel()
    g32 = (g_mine & 0xFFFFFFFF).to(torch.int32)
    for k in range(epa):   # one "round" of events at a time keeps the temporaries small (matters for the 64 GB case)
        blk = r[k * na:(k + 1) * na]
        u = torch.rand(na, generator=gen, device=dev)
        blk[:, 0] = torch.where(u < 0.45, 0, torch.where(u < 0.9, 1, 2)).to(torch.int32)
        blk[:, 1] = k + 1
        blk[:, 2] = g32
        blk[:, 4] = torch.randint(0, 1 << 31, (na,), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
    del u, g32
    # ownership: partition = a multiplicative hash of the dense id (ids are pre-hashed once on load), 64 partitions
    part = ((np.arange(n_global, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)).astype(np.uint32) % np.uint32(64)
    cap = int(n * 1.15) + 1_000_000
    res = {"workload": (f"configs[2] FULL problem, strong: {n_global} aggregates x {epa} events x 64 B = {n_global * epa * 64 / 1e9:.0f} GB split over {world} rank(s)"
                        if strong else f"configs[2] shape, weak: {ROUTED_AGG_PER_GPU} aggregates x {epa} events x 64 B originate per GPU, "
                        f"{n_global} aggregates hash-partitioned over {world} rank(s)"), "events_total": int(n) * world}
    # sort-free: the arrived records are folded with integer atomics (K6 kernel), no group-by; sorted_group: K5 + K1
    for mode, fused, sort_based in (("nccl_all_to_all", False, False), ("fused_peer_scatter", True, False), ("nccl_all_to_all_sorted_group", False, True)):
        if world == 1 and fused:
            continue
        if strong and sort_based:
            continue   # the group-by's scratch does not fit next to 64 GB of records on one GPU
        eng = ReplayEngine(local_rank)
        eng.register_program(P.counter_program())
        if sort_based:
            eng.set_option("incremental", 1)
        D.exchange_ids(eng, rank, world, cap, fused=fused)
        eng.dist_set_partitions(part)
        best = None
        for it in range(3):
            barrier()
            t0 = time.perf_counter()
            eng.dist_route_and_fold(r.view(torch.uint8), fused)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            ds = eng.dist_stats()
            v = [dt, ds.ms_count, ds.ms_counts_exchange, ds.ms_scatter, ds.ms_exchange, ds.ms_group, ds.ms_fold]
            t = torch.tensor(v, dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            v = [float(x) for x in t]
            if it > 0 and (best is None or v[0] < best[0]):
                best = v
        ds = eng.dist_stats()
        ev = int(eng.stats().n_events)
        tot = torch.tensor([ev], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        assert int(tot[0]) == n * world, (int(tot[0]), n * world)
        res[("single_gpu_" + ("sorted_group" if sort_based else "sort_free")) if world == 1 else mode] = {
            "events_per_s": n * world / best[0], "ms_wall": best[0] * 1e3, "ms_route_count": best[1], "ms_counts_exchange": best[2],
            "ms_route_scatter": best[3], "ms_exchange": best[4], "ms_group": best[5], "ms_fold": best[6],
            "fold_events_per_s_per_gpu": ds.n_recv / (best[6] * 1e-3) if best[6] else None,
            "remote_fraction": ds.n_sent_remote / max(ds.n_sent, 1)}
        eng.close()
        del eng
        torch.cuda.empty_cache()
    return res


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer region (default min(steps, 20))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true", help="progress markers on stderr")
    ap.add_argument("--no-routed", action="store_true", help="skip the routed (configs[2]-shaped) pipeline measurement")
    ap.add_argument("--routed-strong", action="store_true",
                    help="routed pipeline on the FULL configs[2] problem (10 M aggregates x 100 events = 64 GB) split over the ranks "
                         "(strong scaling; needs 64 GB of records on one GPU at N=1) instead of 8 GB per rank")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    def note(msg):
        if args.verbose:
            print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)

    note("importing torch")
    import numpy as np
    import torch
    import torch.distributed as dist

    from surge_b200 import ReplayEngine
    from surge_b200 import programs as P
    from surge_b200 import synth as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the replay engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- this rank's shard: its own 2^20 aggregates (hash-partitioned aggregates are independent units)
    note("generating the log on the device")
    rec, off = S.counter_csr_device(N_AGG, EVENTS_PER_AGG, seed=2 + rank, device=dev)
    n_events = N_AGG * EVENTS_PER_AGG
    log_bytes = int(rec.numel() * 4)
    b_alg = algorithmic_bytes(N_AGG, EVENTS_PER_AGG)
    eng = ReplayEngine(local_rank)
    eng.register_program(P.counter_program())
    eng.load_events(rec.view(torch.uint8), off)
    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)

    W = max(args.warmup, 3)
    K = args.steps
    note("warm-up folds")
    for _ in range(W):
        eng.set_initial_states(None)
        eng.fold()
    kernel_ms = eng.stats().ms_fold  # one fold alone, CUDA events around the kernel

    note("timed region 1")
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    # ---- timed region 1 (value): K pipelined folds, inputs resident in HBM
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(K):
        eng.set_initial_states(None)   # every step is a full rebuild from None
        eng.fold_async()
    ev1.record(stream)
    eng.wait()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = K * int(eng.stats().fold_launches)
    folded_events = int(eng.stats().n_events)
    assert folded_events == n_events, (folded_events, n_events)

    # ---- timed region 2 (e2e): host buffers through the C ABI, H2D + fold + D2H every step
    ke = args.e2e_steps or min(K, 20)
    note("pinned host buffers")
    host_log = torch.empty(rec.numel() * 4, dtype=torch.uint8, pin_memory=True)
    host_log.copy_(rec.view(torch.uint8).view(-1))
    host_off = off.cpu().numpy().astype(np.uint64)
    host_states = torch.empty(N_AGG * STATE_BYTES, dtype=torch.uint8, pin_memory=True)
    host_log_np, host_states_np = host_log.numpy(), host_states.numpy().reshape(N_AGG, STATE_BYTES)
    e2 = ReplayEngine(local_rank)
    e2.register_program(P.counter_program())
    for _ in range(2):
        e2.load_events(host_log_np, host_off)
        e2.set_initial_states(None)
        e2.fold()
        e2.export_states(host_states_np)
    barrier()
    note("timed region 2 (e2e)")
    t0 = time.perf_counter()
    for _ in range(ke):
        e2.load_events(host_log_np, host_off)      # H2D of the step's input from pinned host memory
        e2.set_initial_states(None)
        e2.fold()
        e2.export_states(host_states_np)           # D2H of the step's result
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    st2 = e2.stats()
    # the e2e result must be the same table the resident fold produced
    same = bool(torch.equal(torch.from_numpy(host_states_np.reshape(-1)).to(dev), eng.states_tensor().reshape(-1)))
    assert same, "e2e state table differs from the HBM-resident fold"

    # ---- configs[2] shape: events arrive by source partition, one exchange routes them to the owning rank
    routed = None
    if not args.no_routed:
        try:
            note("routed pipeline")
            e2.close(); eng.close()
            del rec
            torch.cuda.empty_cache()
            routed = routed_pipeline(rank, world, local_rank, dev, barrier, note, strong=args.routed_strong)
        except Exception as ex:  # noqa: BLE001 - the headline line must survive a failure of the extra measurement
            routed = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(t[0]), float(t[1])
    value = world * n_events * K / (ms_total * 1e-3)
    e2e_value = world * n_events * ke / e2e_s

    out = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = b_alg / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_fold_runs_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp))["dram_bytes_per_launch"]
            except Exception:  # noqa: BLE001
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "aggregates_per_gpu": N_AGG, "events_per_aggregate": EVENTS_PER_AGG,
                       "record_bytes": 64, "state_bytes": STATE_BYTES, "model": "Counter (scaladsl TestBoundedContext)",
                       "l2": "inputs (2 GiB log per GPU) are 16x the 126 MB L2; no flush between iterations",
                       "sharding": "aggregates sharded across ranks, no data-path collective (see DESIGN.md multi-GPU)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": b_alg, "kernel": "fold_runs_kernel",
                         "kernel_ms": kernel_ms, "peak_source": peak_src,
                         "pipelined_frac": (b_alg / (ms_total / K * 1e-3) / 1e9) / peak},
            "e2e": {"value": e2e_value, "unit": "events/s", "h2d_bytes_per_step": int(log_bytes + host_off.nbytes),
                    "d2h_bytes_per_step": int(N_AGG * STATE_BYTES), "steps": ke, "ms_per_step": e2e_s / ke * 1e3,
                    "ms_h2d": float(st2.ms_h2d), "ms_fold": float(st2.ms_fold), "ms_d2h": float(st2.ms_d2h),
                    # the end-to-end step is the PCIe copy of the log: its rate is the bound of this number, not the kernel
                    "h2d_gb_per_s": (float(log_bytes + host_off.nbytes) / (float(st2.ms_h2d) * 1e-3) / 1e9) if st2.ms_h2d > 0 else None,
                    "h2d_share_of_step": (float(st2.ms_h2d) / (e2e_s / ke * 1e3)) if e2e_s > 0 else None},
            "clocks": clocks,
        }
        if routed is not None:
            out["routed"] = routed
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cpu_rec = np.array(host_log_np, copy=True)   # pageable copy: pinned memory is not what a CPU-only deployment would read
            v, reps, secs = time_cpu_oracle(cpu_rec, host_off, cores, min_seconds=8.0, max_reps=20)
            out["cpu_baseline"] = {"value": v, "unit": "events/s", "cores": cores, "kind": "port",
                                   "sample": f"full configs[1] log ({n_events} events, pageable host memory) x {reps} passes, {secs:.1f} s, oracle/sgr_oracle.c with {cores} threads"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
