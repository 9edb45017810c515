"""Step-by-step version of tests/test_gpu_program_fuzz.py for diagnosing a failing seed: every step reports separately."""
import importlib.util
import os
import sys
import traceback

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("tf", os.path.join(ROOT, "tests", "test_gpu_program_fuzz.py"))
tf = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tf)
from oracle import program_interp as I  # noqa: E402
from surge_b200 import ReplayEngine  # noqa: E402
from surge_b200 import native as N  # noqa: E402
from surge_b200 import programs as P  # noqa: E402


def step(name, fn):
    try:
        fn()
        print("   ok  ", name, flush=True)
        return True
    except Exception as ex:  # noqa: BLE001
        print("   FAIL", name, "->", type(ex).__name__, str(ex)[:600], flush=True)
        return False


for seed in [int(a) for a in sys.argv[1:]] or range(6):
    rng = np.random.default_rng(9000 + seed)
    sb, rules, f64 = tf.draw_program(rng)
    print(f"seed {seed}: state_bytes {sb} rules {rules} f64 {f64}", flush=True)
    prog = P.make_program(sb, N.REC_FIXED64, rules, f64_fields=f64)
    rec, off, aggs = tf.draw_log(rng, len(rules), 260, 700, f64)
    want = I.fold(rules, sb, rec, off, f64_fields=f64)
    try:
        e = ReplayEngine(0)
        if not step("register", lambda: e.register_program(prog)):
            continue
        for k in (0, 1, 3):
            def run(k=k):
                e.set_option("kernel", k)
                e.set_initial_states(None)
                e.load_events(rec, off)
                e.fold()
                tf.same(e.export_states(), want, f"kernel {k}")
            step(f"fold kernel {k}", run)
        rec2, off2, _ = tf.draw_log(rng, len(rules), 260, 300, f64)
        want2 = I.fold(rules, sb, rec2, off2, initial=want, f64_fields=f64)
        for k in (0, 1):
            def run2(k=k):
                e.set_option("kernel", k)
                e.set_initial_states(want)
                e.load_events(rec2, off2)
                e.fold()
                tf.same(e.export_states(), want2, f"prior kernel {k}")
            step(f"fold with prior states kernel {k}", run2)
        e.set_option("kernel", 0)
        perm = tf.interleave(rng, aggs)

        def run3():
            e.set_initial_states(None)
            e.fold_unsorted(rec[perm], 260)
            tf.same(e.export_states(), want, "fold_unsorted")
        step("fold_unsorted", run3)
        table = want
        e.set_initial_states(want)
        for b in range(3):
            recb, _, aggb = tf.draw_log(rng, len(rules), 260, [40, 400, 5][b], f64)
            batch = recb[tf.interleave(rng, aggb)]
            table = I.fold_arrival_order(rules, sb, batch, table, f64_fields=f64)

            def run4(batch=batch, table=table):
                e.fold_incremental(batch)
                tf.same(e.export_states(), table, "micro-batch")
            if not step(f"micro-batch {b}", run4):
                e.set_initial_states(table)
        e.close()
    except Exception:  # noqa: BLE001
        traceback.print_exc()
