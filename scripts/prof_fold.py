"""Profiling target: config 2 (or a slice of it), one warm-up fold and N folds of the chosen kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from surge_b200 import ReplayEngine, programs as P, synth as S
kernel = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_agg = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
folds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
variant = int(sys.argv[4]) if len(sys.argv) > 4 else -1
rec, off = S.counter_csr_device(n_agg, 32, seed=2)
e = ReplayEngine(0); e.register_program(P.counter_program()); e.set_option("kernel", kernel)
if variant >= 0: e.set_option("fold_variant" if kernel == 1 else "run_variant", variant)
e.load_events(rec.view(torch.uint8), off)
for _ in range(1 + folds):
    e.set_initial_states(None); e.fold()
print("ms_fold", e.stats().ms_fold, "events", e.stats().n_events)
