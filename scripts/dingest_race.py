"""Bisect helper for the device ingest: the same 4 M-record poll under the SGR_DINGEST_DEBUG variants, several repetitions each
(one process per run: the flags are read once). Prints one line per run."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from oracle import oracle as O
from surge_b200 import ReplayEngine, programs as P
from surge_b200.dingest import DeviceIngest
n = 4_000_000
rng = np.random.default_rng(1)
agg = rng.integers(0, 200_000, size=n).astype(np.uint32)
wire = O.kafka_encode_counter(agg, rng.integers(0, 3, size=n).astype(np.uint32), np.arange(n, dtype=np.uint32), rng.integers(0, 1 << 31, size=n).astype(np.int32), 512, True)
ok = bad = 0
msgs = []
with ReplayEngine(0) as e:
    e.register_program(P.counter_program())
    with DeviceIngest(e, 1 << 19) as dg:
        for it in range(int(sys.argv[1])):
            e.set_initial_states(None); dg.reset()
            dg.submit(0, wire)
            try:
                dg.fold(); ok += 1
            except Exception as ex:
                bad += 1; msgs.append(str(ex)[-90:])
print("ok", ok, "bad", bad, msgs[:3])
''' % ROOT
reps = sys.argv[1] if len(sys.argv) > 1 else "6"
for flags in sys.argv[2:] or ["0", "16", "1", "2", "4", "8"]:
    env = dict(os.environ, SGR_DINGEST_DEBUG=flags)
    r = subprocess.run([sys.executable, "-c", CHILD, reps], env=env, capture_output=True, text=True, timeout=600)
    print("flags", flags, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)
