"""Dev script: config-2 timing per kernel (device events around K async folds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surge_b200 import ReplayEngine, programs as P, synth as S

def big(n_agg, epa, kernels):
    rec, off = S.counter_csr_device(n_agg, epa, seed=2)
    torch.cuda.synchronize()
    nbytes = rec.numel() * 4
    balg = nbytes + 8 * (n_agg + 1) + 16 * n_agg
    for kernel, variant in kernels:
        e = ReplayEngine(0); e.register_program(P.counter_program()); e.set_option("kernel", kernel)
        if variant is not None: e.set_option("fold_variant" if kernel == 1 else "run_variant", variant)
        e.load_events(rec.view(torch.uint8), off)
        for _ in range(3):
            e.set_initial_states(None); e.fold()
        s = torch.cuda.ExternalStream(e.stream_ptr())
        K = 10
        with torch.cuda.stream(s):
            ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record(s)
            for _ in range(K):
                e.set_initial_states(None); e.fold_async()
            ev1.record(s)
        e.wait(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / K
        print(f"n_agg={n_agg} epa={epa} kernel={kernel} variant={variant}: ms/fold={ms:.4f} last_kernel_ms={e.stats().ms_fold:.4f} "
              f"GB/s={balg/ms/1e6:.0f} Gev/s={n_agg*epa/ms/1e6:.2f} launches={e.stats().fold_launches}", flush=True)
        e.close()

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    ks = [(2, v) for v in range(7)] + [(3, None), (1, 1)]
    big(1 << 20, 32, ks)
    big(10_000_000 // 8, 100, ks)
