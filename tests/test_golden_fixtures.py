"""The committed golden fixtures (tests/golden/appendix_d_vectors.json: the reference's own known-answer vectors for the
replay path) against (a) both CPU restatements of the oracle and (b) the CUDA path through the C ABI."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import surge_model as M
from surge_b200 import formats as F
from surge_b200 import native as N
from surge_b200 import programs as P

V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_d_vectors.json")))
ID = "agg"


def _records(events, agg=0):
    t = {"Incr": F.COUNT_INCREMENTED, "Decr": F.COUNT_DECREMENTED, "NoOp": F.NO_OP_EVENT, "Throw": F.EXCEPTION_THROWING_EVENT}
    rows = [(t[e[0]], e[-1], agg, e[1] if len(e) == 3 else 0) for e in events]
    return F.counter_records([r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows], [r[3] for r in rows])


def _objects(events):
    out = []
    for e in events:
        out.append({"Incr": lambda: M.CountIncremented(ID, e[1], e[2]), "Decr": lambda: M.CountDecremented(ID, e[1], e[2]),
                    "NoOp": lambda: M.NoOpEvent(ID, e[1]), "Throw": lambda: M.ExceptionThrowingEvent(ID, e[1], RuntimeError("failed"))}[e[0]]())
    return out


def _prior_row(prior):
    row = np.zeros(1, dtype=F.COUNTER_STATE)
    if prior is not None:
        row["count"], row["version"], row["flags"] = prior[0], prior[1], N.ST_EXISTS
    return row


def _expect_flags(v):
    fl = N.ST_EXISTS if v["state"] is not None else 0
    if v["error"]:
        return (N.ST_EXISTS if v["prior"] is not None else 0) | N.ST_ERROR
    return fl | (N.ST_CHANGED if v["published"] else 0)


@pytest.mark.parametrize("family,model,handler", [("counter", O.MODEL_COUNTER, M.counter_handle_event), ("ml_counter", O.MODEL_ML_COUNTER, M.ml_counter_apply_event)])
def test_oracle_restatements_match_the_fixtures(family, model, handler):
    for v in V[family]:
        prior = None if v["prior"] is None else M.State(ID, *v["prior"])
        ack = M.apply_events(handler, prior, _objects(v["events"]))
        assert ack.success == (not v["error"]), v["source"]
        assert (None if ack.state is None else [ack.state.count, ack.state.version]) == v["state"], v["source"]
        assert ack.published_state == v["published"], v["source"]
        rec = _records(v["events"])
        out, _, nerr = O.fold_packed(model, O.REC_FIXED64, rec, F.csr_offsets_from_counts([len(rec)]), _prior_row(v["prior"]))
        row = out.view(F.COUNTER_STATE).reshape(-1)[0]
        assert [int(row["count"]), int(row["version"])] == (v["state"] or [0, 0]) and int(row["flags"]) == _expect_flags(v), v["source"]
        assert int(row["err_idx"]) == v.get("err_idx", 0) and nerr == int(v["error"])


def test_ktable_fixture():
    for v in V["ktable"]:
        js = lambda s, i: ('{"string":"%s","int":%d}' % (s, i)).encode()  # noqa: E731
        table = M.ktable_restore([(k, js(k, i)) for k, i in v["records"]])
        assert table == {k: js(k, i) for k, i in v["table"].items()}, v["source"]


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [0, 1, 3])
def test_cuda_path_matches_the_fixtures(kernel):
    from surge_b200 import ReplayEngine

    for family, prog in (("counter", P.counter_program()), ("ml_counter", P.ml_counter_program())):
        vs = V[family]
        rec = np.concatenate([_records(v["events"], i) for i, v in enumerate(vs)])
        off = F.csr_offsets_from_counts([len(v["events"]) for v in vs])
        init = np.concatenate([_prior_row(v["prior"]) for v in vs])
        with ReplayEngine(0) as e:
            e.register_program(prog)
            e.set_option("kernel", kernel)
            e.set_initial_states(init)
            e.load_events(rec, off)
            e.fold()
            rows = e.export_states().view(F.COUNTER_STATE).reshape(-1)
        for i, v in enumerate(vs):
            assert [int(rows[i]["count"]), int(rows[i]["version"])] == (v["state"] or [0, 0]), (kernel, v["source"])
            assert int(rows[i]["flags"]) == _expect_flags(v) and int(rows[i]["err_idx"]) == v.get("err_idx", 0), (kernel, v["source"])


@pytest.mark.gpu
def test_cuda_bank_account_fixture():
    from surge_b200 import ReplayEngine

    for v in V["bank_account"]:
        blob = b""
        for j, ev in enumerate(v["events"]):
            blob += F.bank_created_record(0, j + 1, v["account"], ev[1], ev[2], ev[3]) if ev[0] == "Created" else F.bank_updated_record(0, j + 1, v["account"], ev[1])
        with ReplayEngine(0) as e:
            e.register_program(P.bank_account_program())
            e.load_events(np.frombuffer(blob, np.uint8), F.csr_offsets_from_counts([len(v["events"])]))
            e.fold()
            st = F.decode_bank_state(e.export_states().view(F.BANK_STATE).reshape(-1)[0])
        assert st == dict(v["state"], accountNumber=v["account"]), v["source"]
        out, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, np.frombuffer(blob, np.uint8), F.csr_offsets_from_counts([len(v["events"])]))
        assert F.decode_bank_state(out.view(F.BANK_STATE).reshape(-1)[0]) == st
