"""configs[3] (Zipf alpha=1.1 keys, variable payloads 32-512 B), scaled: n_keys and n_events are arguments.
Builds the log on the GPU with torch (headers + directory), folds it with the record-parallel variable kernel
(and optionally the lane-per-aggregate TMA kernel for comparison)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surge_b200 import ReplayEngine, programs as P, native as N, synth as S
n_keys = int(sys.argv[1]) if len(sys.argv) > 1 else 625_000
n_events = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
compare = len(sys.argv) > 3 and sys.argv[3] == "compare"
stages = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [12288]
nst = int(sys.argv[5]) if len(sys.argv) > 5 else 2
dev = "cuda:0"
counts_np = S.zipf_counts(n_keys, n_events, 1.1, seed=4)
counts = torch.from_numpy(counts_np).to(dev)
gen = torch.Generator(device=dev); gen.manual_seed(4)
plen = torch.randint(32, 513, (n_events,), generator=gen, device=dev, dtype=torch.int64)
rlen = 16 + ((plen + 15) // 16) * 16
rec_off = torch.zeros(n_events + 1, dtype=torch.int64, device=dev); rec_off[1:] = torch.cumsum(rlen, 0)
total = int(rec_off[-1])
starts = torch.zeros(n_keys + 1, dtype=torch.int64, device=dev); starts[1:] = torch.cumsum(counts, 0)
seg = rec_off[starts]
buf = torch.randint(0, 256, (total,), generator=gen, device=dev, dtype=torch.uint8)   # filler the fold must still read
agg = torch.repeat_interleave(torch.arange(n_keys, device=dev, dtype=torch.int64), counts)
seq = (torch.arange(n_events, device=dev, dtype=torch.int64) - starts[agg] + 1)
u = torch.rand(n_events, generator=gen, device=dev)
typ = torch.where(u < 0.45, 0, torch.where(u < 0.9, 1, 2))
hdr = torch.stack([typ, seq, plen, agg], 1).to(torch.int32)                            # type, seq, payload_len, agg
w = buf.view(torch.int32)
pos = (rec_off[:-1] // 4)
for k in range(4): w[pos + k] = hdr[:, k]
del u, typ, seq, agg, hdr, pos, plen, rlen
torch.cuda.synchronize()
b_alg = total + 8 * (n_keys + 1) + 16 * n_keys + 8 * (n_events + 1)
hot = int(counts_np.max())
res = {"config": "configs[3] scaled", "keys": n_keys, "events": n_events, "log_GiB": total / 2**30, "hottest_key_share": hot / n_events}
for name, kernel, sb in ([(f"vruns stage={sb}", 0, sb) for sb in stages] + ([("lane-per-aggregate TMA", 1, 0)] if compare else [])):
    e = ReplayEngine(0); e.register_program(P.counter_program(N.REC_VAR16)); e.set_option("kernel", kernel)
    if sb: e.set_option("var_stage_bytes", sb); e.set_option("var_stages", nst)
    e.load_events_indexed(buf, seg, rec_off)
    ms = []
    for _ in range(3):
        e.set_initial_states(None); e.fold(); ms.append(e.stats().ms_fold)
    st = e.stats()
    res[name] = {"ms_fold": min(ms), "events_per_s": n_events / (min(ms) * 1e-3), "GBps": b_alg / (min(ms) * 1e-3) / 1e9, "n_events": int(st.n_events),
                 "launches": int(st.fold_launches), "errors": int(st.n_errors)}
    if kernel == 0 and sb == stages[0]: ref = e.states_tensor().clone()
    else: res[name]["table_equals_first"] = bool(torch.equal(ref, e.states_tensor()))
    e.close()
print(json.dumps(res))
