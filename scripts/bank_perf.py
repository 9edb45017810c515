"""BankAccount (64-byte state, IF_EXISTS rule) on the configs[1] shape: runs kernel (W=14) vs lane-sequential TMA kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surge_b200 import ReplayEngine, programs as P
n_agg, epa = 1 << 20, 32
dev = "cuda:0"; n = n_agg * epa
gen = torch.Generator(device=dev); gen.manual_seed(9)
r = torch.randint(-(1 << 31), 1 << 31, (n, 16), generator=gen, device=dev, dtype=torch.int64).to(torch.int32)
idx = torch.arange(n, device=dev, dtype=torch.int64)
r[:, 0] = (idx % epa != 0).to(torch.int32)          # first event of every account creates it, the rest update the balance
r[:, 1] = (idx % epa + 1).to(torch.int32)
r[:, 2] = (idx // epa).to(torch.int32); r[:, 3] = 0
off = torch.arange(n_agg + 1, device=dev, dtype=torch.int64) * (epa * 64)
b_alg = n * 64 + 8 * (n_agg + 1) + 64 * n_agg
tabs = []
for kernel in (0, 1):
    e = ReplayEngine(0); e.register_program(P.bank_account_program()); e.set_option("kernel", kernel)
    e.load_events(r.view(torch.uint8), off)
    ms = []
    for _ in range(4):
        e.set_initial_states(None); e.fold(); ms.append(e.stats().ms_fold)
    tabs.append(e.states_tensor().clone())
    print(f"BankAccount kernel={kernel}: {min(ms):.4f} ms  {b_alg/min(ms)/1e6:.0f} GB/s  {n/min(ms)/1e6:.2f} G events/s  launches={e.stats().fold_launches}", flush=True)
    e.close()
print("tables equal:", bool(torch.equal(tabs[0], tabs[1])))
