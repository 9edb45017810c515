"""configs[4] (streaming micro-batch): 100k-event batches appended to 1,048,576 live aggregates, incremental fold."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from surge_b200 import ReplayEngine, programs as P, synth as S
n_agg = 1 << 20; batch = 100_000
n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
check = len(sys.argv) > 2 and sys.argv[2] == "check"
dev = "cuda:0"
rec, off = S.counter_csr_device(n_agg, 4, seed=5, device=dev)
e = ReplayEngine(0); e.register_program(P.counter_program())
e.load_events(rec.view(torch.uint8), off); e.fold()
gen = torch.Generator(device=dev); gen.manual_seed(55)
nb_pool = min(n_batches, 200)   # 200 distinct batches (1.28 GB), cycled
pool = torch.zeros((nb_pool, batch, 16), dtype=torch.int32, device=dev)
u = torch.rand((nb_pool, batch), generator=gen, device=dev)
pool[:, :, 0] = torch.where(u < 0.45, 0, torch.where(u < 0.9, 1, 2)).to(torch.int32)
pool[:, :, 1] = torch.arange(batch, device=dev, dtype=torch.int32)[None, :]
pool[:, :, 2] = torch.randint(0, n_agg, (nb_pool, batch), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
pool[:, :, 4] = torch.randint(0, 1 << 31, (nb_pool, batch), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
torch.cuda.synchronize()
if check:
    from oracle import oracle as O
    want = e.export_states()
    for b in range(5):
        want = O.fold_incremental(O.MODEL_COUNTER, pool[b].cpu().numpy().view(np.uint8).reshape(-1), want)
        e.fold_incremental(pool[b].view(torch.uint8))
    print("parity after 5 batches:", bool(np.array_equal(e.export_states(), want)))
print("warm", flush=True)
for b in range(10): e.fold_incremental(pool[b % nb_pool].view(torch.uint8))
print("warm done", e.stats().ms_fold, flush=True)
lat = []
torch.cuda.synchronize(); t0 = time.perf_counter()
for b in range(n_batches):
    t1 = time.perf_counter()
    e.fold_incremental(pool[b % nb_pool].view(torch.uint8))
    lat.append(time.perf_counter() - t1)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
st = e.stats(); lat = np.array(lat) * 1e6
print(json.dumps({"config": "configs[4] streaming micro-batch", "batches": n_batches, "events_per_batch": batch, "live_aggregates": n_agg,
                  "events_per_s": n_batches * batch / dt, "batch_latency_us": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)), "mean": float(lat.mean())},
                  "last_batch": {"ms_group": st.ms_group, "ms_fold": st.ms_fold, "touched": int(st.n_aggregates), "launches": int(st.fold_launches)}}))
