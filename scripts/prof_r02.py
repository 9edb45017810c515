"""Workload for the round-2 ncu captures (run under ncu with a kernel filter): the per-GPU shape of configs[2] on 8 ranks
(1.25 M aggregates x 100 events, arrival order) through (1) the bulk sort-free fold, (2) the partition + fold pipeline on one
rank (force_route: every region local), (3) one device-ingest poll of 4 M lz4 records."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import oracle as O
from surge_b200 import ReplayEngine
from surge_b200 import programs as P
from surge_b200 import synth as S
from surge_b200.dingest import DeviceIngest

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n_agg, epa = 1_250_000, 100
dev = "cuda:0"
if what in ("all", "bulk", "push"):
    rec = S.routed_log_device(0, 1, n_agg, epa, 3, dev)
    flat = rec.view(torch.uint8).view(-1)
if what in ("all", "bulk"):
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        for _ in range(3):
            e.fold_unsorted(flat, n_agg)
        print("bulk", e.stats().ms_fold, "ms", hex(e.states_hash()))
if what in ("all", "push"):
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        e.set_option("force_route", 1)
        e.dist_init(0, 1, None, rec.shape[0] + 16 * 1024)
        e.dist_set_partitions(np.zeros(n_agg, dtype=np.uint32))
        for _ in range(3):
            e.dist_route_and_fold(flat, 2)
        print("push", e.dist_stats().ms_pipeline, "ms", hex(e.states_hash()))
if what in ("all", "dingest"):
    n = 4_000_000
    rng = np.random.default_rng(1)
    agg = rng.integers(0, 200_000, size=n).astype(np.uint32)
    wire = O.kafka_encode_counter(agg, rng.integers(0, 3, size=n).astype(np.uint32), np.arange(n, dtype=np.uint32), rng.integers(0, 1 << 31, size=n).astype(np.int32), 512, True)
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        with DeviceIngest(e, 1 << 19) as dg:
            for _ in range(2):
                e.set_initial_states(None); dg.reset()
                dg.submit(0, wire)
                st = dg.fold()
            print("dingest", st["n_records"], dg.last_timing())
