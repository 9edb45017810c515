"""A/B of the pipelined push (route_push.cu) under torchrun (or alone on one GPU with force_route): tile size, chunk count,
full vs projected records, on the configs[2] log at a given scale. Prints device pipeline time (max over ranks), the NVLink rate
and the state hash (must not change between variants). Measurement only; parity lives in tests/ and bench.py."""
import os
import sys

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
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["512:16:2:1:2:0", "1024:16:2:1:2:0", "256:16:2:1:2:0", "256:16:2:1:1:0", "512:16:2:0:2:0", "512:16:3:1:2:0", "256:16:3:1:2:0", "512:16:3:0:2:0"]
    n_global = int(10_000_000 * scale) // 64 * 64
    rec = S.routed_log_device(rank, world, n_global, 100, 3, dev)
    flat = rec.view(torch.uint8).view(-1)
    n = rec.shape[0]
    part = S.routed_partitions(n_global, 64)
    e = ReplayEngine(lr)
    e.register_program(P.counter_program())
    if world == 1:
        e.set_option("force_route", 1)
    D.exchange_ids(e, rank, world, int(n * 1.12) + 64 * 1024 * world * 4, fused=True)
    e.dist_set_partitions(part)
    hashes = set()
    for v in variants:
        tile, chunks, fused, pull, fbs, staged = (int(x) for x in v.split(":"))
        e.set_option("push_staged", staged)
        e.set_option("push_fold_blocks_per_sm", fbs)
        e.set_option("push_tile", tile)
        e.set_option("push_pull", pull)
        e.set_option("push_chunks", chunks)
        best = None
        for it in range(4):
            if world > 1:
                dist.barrier()
            e.dist_route_and_fold(flat, fused)
            ds = e.dist_stats()
            t = torch.tensor([ds.ms_pipeline, ds.ms_scatter], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it > 0 and (best is None or float(t[0]) < best[0]):
                best = (float(t[0]), float(t[1]))
        h = torch.tensor([np.int64(np.uint64(e.states_hash()))], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(h)
        hashes.add(int(h.item()))
        ds = e.dist_stats()
        if rank == 0:
            wire = ds.exchange_record_bytes
            print(f"tile={tile:5d} chunks={chunks:3d} fused={fused} pull={pull} fold_blocks/SM={fbs} staged={staged}: pipeline {best[0]:8.3f} ms (push issue {best[1]:8.3f})  {n_global * 100 / best[0] / 1e6:7.2f} G ev/s job  "
                  f"NVLink out {ds.n_sent_remote * wire / best[0] / 1e6:6.1f} GB/s/GPU  source read {n * 64 / best[0] / 1e6:6.1f} GB/s/GPU  hash {int(h.item()) & ((1 << 64) - 1):016x}", flush=True)
    assert len(hashes) == 1, hashes
    e.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
