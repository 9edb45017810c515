"""A/B of the sort-free arrival-order fold (one GPU): the per-GPU shape of configs[2] on 8 ranks — 1.25 M aggregates x 100
events = 125 M records (8 GB) in round-robin arrival order — through
  * the single-launch micro-batch kernel (incremental.cu, round 1),
  * the bulk kernels (bulk_fold.cu) with the cache-hint / unroll / grid knobs,
  * the push pipeline on one rank (force_route): partition kernel + chunked fold with everything local.
Prints ms (CUDA events inside the engine), TB/s of algorithmic bytes and the state hash of every variant (all must agree).
No oracle here: parity of each variant is covered by tests/test_gpu_dist.py; this only measures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from surge_b200 import ReplayEngine
from surge_b200 import programs as P
from surge_b200 import synth as S


def main():
    n_agg = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
    epa = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    dev = "cuda:0"
    rec = S.routed_log_device(0, 1, n_agg, epa, 3, dev)
    n = rec.shape[0]
    b_alg = n * 64 + 48 * n_agg
    flat = rec.view(torch.uint8).view(-1)
    hashes = {}

    def run(label, opts, reps=5):
        with ReplayEngine(0) as e:
            e.register_program(P.counter_program())
            for k, v in opts.items():
                e.set_option(k, v)
            ms = []
            for _ in range(reps):
                e.fold_unsorted(flat, n_agg)
                ms.append(e.stats().ms_fold)
            h = e.states_hash()
        best = min(ms[1:])
        hashes[label] = h
        print(f"{label:46s} {best:8.3f} ms  {n / best / 1e6:7.2f} G ev/s  {b_alg / best / 1e9:6.3f} TB/s  hash {h:016x}", flush=True)

    run("micro-batch kernel (round 1)", {"bulk": 0})
    for hints in (1, 0):
        for unroll in (4, 2, 1):
            for bps in (8, 4, 3):
                run(f"bulk hints={hints} unroll={unroll} blocks/SM={bps}", {"bulk_hints": hints, "bulk_unroll": unroll, "bulk_blocks_per_sm": bps})
    # restore defaults (the knobs are process-wide)
    with ReplayEngine(0) as e:
        e.set_option("bulk_hints", 1); e.set_option("bulk_unroll", 4); e.set_option("bulk_blocks_per_sm", 8)
    assert len(set(hashes.values())) == 1, hashes
    # push pipeline, one rank: every record is "sent" to this rank's own receive region
    for fused, chunks in ((2, 16), (2, 4), (3, 16)):
        with ReplayEngine(0) as e:
            e.register_program(P.counter_program())
            e.set_option("force_route", 1)
            e.set_option("push_chunks", chunks)
            e.dist_init(0, 1, None, n + chunks * 1024)
            e.dist_set_partitions(np.zeros(n_agg, dtype=np.uint32))
            best = None
            for _ in range(4):
                e.dist_route_and_fold(flat, fused)
                ds = e.dist_stats()
                if best is None or ds.ms_pipeline < best[0]:
                    best = (ds.ms_pipeline, ds.ms_scatter, ds.ms_fold)
            h = e.states_hash()
        print(f"push fused={fused} chunks={chunks}: pipeline {best[0]:.3f} ms (push {best[1]:.3f}, fold tail {best[2]:.3f})  "
              f"{n / best[0] / 1e6:.2f} G ev/s  hash {h:016x}", flush=True)
        assert h == next(iter(hashes.values())), "push pipeline hash differs"
    # K5 + K1 (what wide / class-1 programs pay on an arrival-order log)
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        e.set_option("incremental", 1)
        for _ in range(2):
            e.fold_unsorted(flat, n_agg)
        st = e.stats()
        print(f"sort-based: group {st.ms_group:.3f} ms + fold {st.ms_fold:.3f} ms  hash {e.states_hash():016x}", flush=True)


if __name__ == "__main__":
    main()
