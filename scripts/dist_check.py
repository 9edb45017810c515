"""Multi-GPU parity check + timing (run under torchrun): the configs[2] shape, scaled down so the CPU oracle finishes in seconds.
Every rank builds the same global log, keeps the records of ITS source partitions in arrival order, then routes + folds
in all four exchange modes (0 NCCL all-to-all, 1 peer scatter, 2 pipelined push, 3 pushed projection) and compares its local
table with the oracle's global result; the sum of the ranks' state hashes must equal the hash of the oracle's table.

  torchrun --nproc-per-node N scripts/dist_check.py [n_global] [events_per_aggregate]

Uses the oracle as the checker only (tests/test_gpu_dist.py runs this script)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from surge_b200 import ReplayEngine
from surge_b200 import dist as D
from surge_b200 import programs as P
from surge_b200 import synth as S


def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = f"cuda:{lr}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    n_global = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    epa = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    n_src_partitions = 64
    rng = np.random.default_rng(7)
    # ownership: the real string hash of synthetic ids (parity with KafkaPartitioner), 32 state-topic partitions
    part = D.partitions_for_keys([f"agg-{g}" for g in range(n_global)], 32)
    from oracle import oracle as O

    counts = rng.integers(max(0, epa - 5), epa + 6, size=n_global)
    rec, off = S.counter_csr(n_global, counts, seed=3, p_throw=0.0005)
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, threads=8)
    want_hash = D.states_hash(want)
    arrival = S.interleave_arrival(rec, seed=4)
    src_part = (arrival["agg"] % n_src_partitions).astype(np.int64)
    mine = arrival[(src_part % world) == rank]
    local_rec = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy()).to(dev)
    cap = int(n_global * epa / world * 1.4) + 16 * 1024 * world * 8
    all_ok = True
    for fused in ([0, 1, 2, 3] if world > 1 else [0, 2, 3]):
        e = ReplayEngine(lr)
        e.register_program(P.counter_program())
        if world == 1:
            e.set_option("force_route", 1)
        e.set_option("push_chunks", 8)
        D.exchange_ids(e, rank, world, cap, fused=fused >= 1)
        e.dist_set_partitions(part)
        dt = 0.0
        for _ in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            e.dist_route_and_fold(local_rec, fused)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ds = e.dist_stats()
        got = e.export_states(); gl = e.dist_local_aggregates()
        ok = bool(np.array_equal(got, want[gl.astype(np.int64)]))
        h = torch.tensor([np.int64(np.uint64(e.states_hash()))], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(h)   # int64 wrap-around sum == sum mod 2^64
        hash_ok = int(np.uint64(np.int64(h.item()))) == want_hash
        all_ok &= ok and hash_ok
        print(f"[rank {rank}/{world}] fused={fused} parity={ok} hash_ok={hash_ok} wall={dt * 1e3:.2f} ms sent={ds.n_sent} remote={ds.n_sent_remote} "
              f"recv={ds.n_recv} n_local={ds.n_local_aggregates} ms: count={ds.ms_count:.3f} scatter/push={ds.ms_scatter:.3f} "
              f"xchg={ds.ms_exchange:.3f} group={ds.ms_group:.3f} fold={ds.ms_fold:.3f} pipeline={ds.ms_pipeline:.3f} bytes/rec={ds.exchange_record_bytes}", flush=True)
        e.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    if not all_ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
