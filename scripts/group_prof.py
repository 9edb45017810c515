"""K5 at scale on one GPU: arrival-order load (stable group-by) + fold, stage times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surge_b200 import ReplayEngine, programs as P
n_agg = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
epa = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = "cuda:0"
n = n_agg * epa
gen = torch.Generator(device=dev); gen.manual_seed(1)
r = torch.zeros((n, 16), dtype=torch.int32, device=dev)
agg = torch.arange(n_agg, device=dev, dtype=torch.int64).repeat(epa)   # event e of every aggregate before event e+1
u = torch.rand(n, generator=gen, device=dev)
r[:, 0] = torch.where(u < 0.45, 0, torch.where(u < 0.9, 1, 2)).to(torch.int32)
r[:, 1] = torch.arange(epa, device=dev, dtype=torch.int32).repeat_interleave(n_agg) + 1
r[:, 2] = agg.to(torch.int32)
r[:, 4] = torch.randint(0, 1 << 31, (n,), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
del agg, u
e = ReplayEngine(0); e.register_program(P.counter_program())
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e.load_unsorted(r.view(torch.uint8), n_agg)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    e.set_initial_states(None); e.fold()
    st = e.stats()
    print(f"n={n} records ({n*64/2**30:.2f} GiB) n_agg={n_agg}: group {st.ms_group:.3f} ms ({n*64*2/st.ms_group/1e6:.0f} GB/s of 2x record bytes) "
          f"fold {st.ms_fold:.3f} ms wall_load {1e3*(t1-t0):.2f} ms events={st.n_events}", flush=True)
