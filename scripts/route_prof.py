"""Profiling target: K4 (route count + scatter) and the sort-free fold on one GPU, configs[2] per-GPU shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from surge_b200 import ReplayEngine, programs as P
n_agg, epa = 1_250_000, int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = "cuda:0"; n = n_agg * epa
gen = torch.Generator(device=dev); gen.manual_seed(1)
r = torch.zeros((n, 16), dtype=torch.int32, device=dev)
u = torch.rand(n, generator=gen, device=dev)
r[:, 0] = torch.where(u < 0.45, 0, torch.where(u < 0.9, 1, 2)).to(torch.int32); del u
r[:, 1] = torch.arange(epa, device=dev, dtype=torch.int32).repeat_interleave(n_agg) + 1
r[:, 2] = torch.arange(n_agg, device=dev, dtype=torch.int64).repeat(epa).to(torch.int32)
r[:, 4] = torch.randint(0, 1 << 31, (n,), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
part = ((np.arange(n_agg, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)).astype(np.uint32) % np.uint32(64)
e = ReplayEngine(0); e.register_program(P.counter_program()); e.set_option("force_route", 1)
e.dist_init(0, 1, None, n + 1000); e.dist_set_partitions(part)
for _ in range(3):
    e.dist_route_and_fold(r.view(torch.uint8), False)
ds = e.dist_stats()
print(f"count={ds.ms_count:.3f} scatter={ds.ms_scatter:.3f} fold={ds.ms_fold:.3f} ms for {n} records")
