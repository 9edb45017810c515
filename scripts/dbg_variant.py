import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
from oracle import oracle as O
from surge_b200 import ReplayEngine, SgrError, dist as D, native as N, programs as P, synth as S
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_dist as T

R, n_global = 4, 20000
rng = np.random.default_rng(1 * 7 + 1 * 3 + 512)
counts = rng.integers(0, 35, size=n_global)
rec, off = S.counter_csr(n_global, counts, seed=512 + 1, p_throw=0.001)
want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
part = D.partitions_for_keys([f"agg-{g}" for g in range(n_global)], 32)
knobs = ReplayEngine(0)
for pull, staged, ordered_first in ((1, 1, 0), (1, 0, 0), (0, 1, 0), (1, 1, 1), (1, 0, 1)):
    knobs.set_option("push_pull", pull); knobs.set_option("push_staged", staged); knobs.set_option("push_tile", 512)
    for fused in (2, 3):
        engines, errors = T._loopback_job(R, n_global, rec, part, fused, chunks=3, capacity=int(len(rec) / R * 1.6) + 200000)
        bad_total = []
        for r, e in enumerate(engines):
            got = e.export_states(); gl = e.dist_local_aggregates().astype(np.int64)
            bad = np.nonzero((got != want[gl]).any(axis=1))[0]
            for b in bad:
                bad_total.append((r, int(gl[b]), got[b].view(np.int32)[:3].tolist(), want[gl[b]].view(np.int32)[:3].tolist()))
        print(f"pull={pull} staged={staged} fused={fused}: errors={[str(x)[:40] if x else None for x in errors]} bad={bad_total}", flush=True)
        if bad_total and fused == 2:
            arrival = S.interleave_arrival(rec, seed=4)
            src = (arrival["agg"] % 64).astype(np.int64) % R
            for (r, g, _, _) in bad_total[:2]:
                s = int((g % 64) % R)
                feed = arrival[src == s]
                pos = np.nonzero(feed["agg"] == g)[0]
                n = len(feed); chunk = ((n + 2) // 3 + 1023) // 1024 * 1024
                print("   agg", g, "source", s, "feed len", n, "chunk_recs", chunk, "positions", pos.tolist(), "types", feed["type"][pos].tolist(), "seq", feed["seq"][pos].tolist(),
                      "chunk/tile", [(int(p // chunk), int((p % chunk) // 512)) for p in pos])
        for e in engines: e.close()
