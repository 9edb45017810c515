"""Multi-GPU check + timing (run under torchrun): config-3 shape, scaled. Every rank builds the same global log,
keeps the records of ITS source partitions in arrival order, routes, folds, and compares its local table with
the oracle's global result."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from surge_b200 import ReplayEngine, programs as P, synth as S
from surge_b200 import dist as D

def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
    n_global = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    epa = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    check = (sys.argv[3] != "nocheck") if len(sys.argv) > 3 else True
    n_src_partitions = 64
    rng = np.random.default_rng(7)
    # ownership: the real string hash of synthetic ids (parity with KafkaPartitioner), 32 state-topic partitions
    if n_global <= 2_000_000:
        part = D.partitions_for_keys([f"agg-{g}" for g in range(n_global)], 32)
    else:
        part = (np.arange(n_global, dtype=np.uint64) * np.uint64(2654435761) >> np.uint64(7)).astype(np.uint32) % 32
    if check:
        from oracle import oracle as O
        counts = rng.integers(max(0, epa - 5), epa + 6, size=n_global)
        rec, off = S.counter_csr(n_global, counts, seed=3, p_throw=0.0005)
        want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, threads=8)
        arrival = S.interleave_arrival(rec, seed=4)
        src_part = (arrival["agg"] % n_src_partitions).astype(np.int64)
        mine = arrival[(src_part % world) == rank]
        local_rec = torch.from_numpy(mine.view(np.uint8).reshape(-1)).to(f"cuda:{lr}")
    else:
        # big: build this rank's share on the device; aggregates of source partition p are those with g % 64 == p
        g_mine = torch.arange(rank, n_global, world, device=f"cuda:{lr}", dtype=torch.int64)  # src partition % world == rank (64 % world == 0)
        n = g_mine.numel() * epa
        r = torch.zeros((n, 16), dtype=torch.int32, device=f"cuda:{lr}")
        gen = torch.Generator(device=f"cuda:{lr}"); gen.manual_seed(100 + rank)
        # arrival order: event e of every aggregate before event e+1 (aggregates interleaved, per-aggregate order kept)
        agg = g_mine.repeat(epa)
        u = torch.rand(n, generator=gen, device=r.device)
        r[:, 0] = torch.where(u < 0.45, 0, torch.where(u < 0.9, 1, 2)).to(torch.int32)
        r[:, 1] = torch.arange(epa, device=r.device, dtype=torch.int32).repeat_interleave(g_mine.numel()) + 1
        r[:, 2] = (agg & 0xFFFFFFFF).to(torch.int32); r[:, 3] = (agg >> 32).to(torch.int32)
        r[:, 4] = torch.randint(0, 1 << 31, (n,), generator=gen, device=r.device, dtype=torch.int64).to(torch.int32)
        local_rec = r.view(torch.uint8).view(-1)
    n_local_rec = local_rec.numel() // 64
    cap = int(n_global * epa / world * 1.3) + 100000
    for fused in ([False, True] if world > 1 else [False]):
        e = ReplayEngine(lr); e.register_program(P.counter_program())
        D.exchange_ids(e, rank, world, cap, fused=fused)
        e.dist_set_partitions(part)
        for it in range(3):
            if world > 1: dist.barrier()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            e.dist_route_and_fold(local_rec, fused)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=f"cuda:{lr}", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t[0])
        ds = e.dist_stats()
        ok = "n/a"
        if check:
            got = e.export_states(); gl = e.dist_local_aggregates()
            ok = bool(np.array_equal(got, want[gl.astype(np.int64)]))
        tot_events = n_global * epa
        print(f"[rank {rank}/{world}] fused={fused} parity={ok} wall={dt*1e3:.2f} ms  {tot_events/dt/1e9:.2f} G events/s (job)  sent={ds.n_sent} remote={ds.n_sent_remote} recv={ds.n_recv} "
              f"n_local={ds.n_local_aggregates} ms: count={ds.ms_count:.3f} cx={ds.ms_counts_exchange:.3f} scatter={ds.ms_scatter:.3f} "
              f"xchg={ds.ms_exchange:.3f} group={ds.ms_group:.3f} fold={ds.ms_fold:.3f}", flush=True)
        e.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    main()
