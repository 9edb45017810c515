"""Property test (hypothesis): the packed C oracle and the object-level restatement agree on arbitrary small Counter logs,
including wrap-around, throws anywhere, prior states and empty batches."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import oracle as O
from oracle import surge_model as M
from surge_b200 import formats as F

AGG = "a"
event = st.one_of(
    st.tuples(st.just(0), st.integers(-2**31, 2**31 - 1), st.integers(0, 2**31 - 1)),
    st.tuples(st.just(1), st.integers(-2**31, 2**31 - 1), st.integers(0, 2**31 - 1)),
    st.tuples(st.just(2), st.just(0), st.integers(0, 2**31 - 1)),
    st.tuples(st.just(3), st.just(0), st.integers(0, 2**31 - 1)),
    st.tuples(st.integers(4, 40), st.just(0), st.just(0)),          # unknown classes: scala.MatchError
)
prior = st.one_of(st.none(), st.tuples(st.integers(-2**31, 2**31 - 1), st.integers(-2**31, 2**31 - 1)))


def to_obj(t, by, seq):
    if t == 0:
        return M.CountIncremented(AGG, by, seq)
    if t == 1:
        return M.CountDecremented(AGG, by, seq)
    if t == 2:
        return M.NoOpEvent(AGG, seq)
    if t == 3:
        return M.ExceptionThrowingEvent(AGG, seq, RuntimeError("x"))
    return object()  # falls to the MatchError branch


@settings(max_examples=300, deadline=None, derandomize=True)
@given(st.lists(event, max_size=12), prior)
def test_c_oracle_equals_object_model(events, pr):
    state = None if pr is None else M.State(AGG, pr[0], pr[1])
    ack = M.apply_events(M.counter_handle_event, state, [to_obj(*e) for e in events])
    rec = F.counter_records([e[0] for e in events], [e[2] for e in events], [0] * len(events), [np.int32(e[1]) for e in events])
    init = np.zeros(1, dtype=F.COUNTER_STATE)
    if pr is not None:
        init["count"], init["version"], init["flags"] = pr[0], pr[1], O.ST_EXISTS
    out, nev, nerr = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, F.csr_offsets_from_counts([len(events)]), init)
    row = out.view(F.COUNTER_STATE).reshape(-1)[0]
    flags = int(row["flags"])
    assert bool(flags & O.ST_ERROR) == (not ack.success) and nerr == int(not ack.success)
    assert bool(flags & O.ST_EXISTS) == (ack.state is not None)
    assert bool(flags & O.ST_CHANGED) == ack.published_state
    if ack.state is not None:
        assert (int(row["count"]), int(row["version"])) == (ack.state.count, ack.state.version)
    if not ack.success:
        first_bad = next(i for i, e in enumerate(events) if e[0] >= 3)
        assert int(row["err_idx"]) == first_bad == nev


# ------------------------------------------------------------------ the program interpreter against the C oracle
# oracle/program_interp.py is the checker of tests/test_gpu_program_fuzz.py; here it is pinned to the C restatement of the
# reference's handlers on the programs that describe them (surge_b200/programs.py), over arbitrary small logs.
import struct  # noqa: E402
import uuid  # noqa: E402

from oracle import program_interp as I  # noqa: E402
from surge_b200 import native as N  # noqa: E402
from surge_b200 import programs as P  # noqa: E402


def rules_of(prog):
    return [(int(prog.rules[t].exists_rule), [(int(o.opcode), int(o.dst_off), int(o.src_off), int(o.len)) for o in prog.rules[t].ops[: prog.rules[t].n_ops]])
            for t in range(prog.n_types)]


@settings(max_examples=200, deadline=None, derandomize=True)
@given(st.lists(st.lists(event, max_size=8), min_size=1, max_size=5), st.lists(prior, min_size=5, max_size=5), st.sampled_from(["counter", "ml_counter", "int_balance"]))
def test_program_interpreter_equals_c_oracle_on_counter_family(segments, priors, which):
    model, prog = {"counter": (O.MODEL_COUNTER, P.counter_program()), "ml_counter": (O.MODEL_ML_COUNTER, P.ml_counter_program()),
                   "int_balance": (O.MODEL_INT_BALANCE, P.int_balance_program())}[which]
    flat = [e for seg in segments for e in seg]
    rec = F.counter_records([e[0] for e in flat], [e[2] for e in flat], [0] * len(flat), [np.int32(e[1]) for e in flat])
    off = F.csr_offsets_from_counts([len(s) for s in segments])
    init = np.zeros(len(segments), dtype=F.COUNTER_STATE)
    for i, pr in enumerate(priors[: len(segments)]):
        if pr is not None:
            init["count"][i], init["version"][i], init["flags"][i] = pr[0], (pr[1] if which != "int_balance" else 0), O.ST_EXISTS
    want, _, _ = O.fold_packed(model, O.REC_FIXED64, rec, off, init)
    got = I.fold(rules_of(prog), 16, rec.view(np.uint8).reshape(-1, 64), off, init.view(np.uint8).reshape(-1, 16))
    assert np.array_equal(got, want.view(np.uint8).reshape(got.shape))


f64s = st.sampled_from([0.0, -0.0, float("nan"), 1.0, -1.0, float("inf"), 5e-324, 1e300])
bank_event = st.one_of(st.tuples(st.just(0), f64s), st.tuples(st.just(1), f64s), st.tuples(st.integers(2, 9), f64s))


@settings(max_examples=200, deadline=None, derandomize=True)
@given(st.lists(st.lists(bank_event, max_size=6), min_size=1, max_size=4), st.lists(st.one_of(st.none(), f64s), min_size=4, max_size=4))
def test_program_interpreter_equals_c_oracle_on_bank_account(segments, priors):
    """Doubles incl. NaN and signed zeros: the publish rule (== on Double after the `eq` shortcut) must agree."""
    ids = [str(uuid.UUID(int=77 + i)) for i in range(len(segments))]
    recs = []
    for i, seg in enumerate(segments):
        for k, (t, bal) in enumerate(seg):
            if t == 0:
                recs.append(F.bank_created_record(i, k + 1, ids[i], "own", "cd", bal))
            else:
                r = bytearray(F.bank_updated_record(i, k + 1, ids[i], bal))
                struct.pack_into("<I", r, 0, t)
                recs.append(bytes(r))
    ev = np.frombuffer(b"".join(recs), np.uint8) if recs else np.zeros(0, np.uint8)
    off = F.csr_offsets_from_counts([len(s) for s in segments])
    init = None
    if any(p is not None for p in priors[: len(segments)]):
        created = b"".join(F.bank_created_record(i, 1, ids[i], "own", "cd", p if p is not None else 0.0) for i, p in enumerate(priors[: len(segments)]))
        init, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, np.frombuffer(created, np.uint8), np.arange(len(segments) + 1, dtype=np.uint64) * 64)
        init = init.copy()
        for i, p in enumerate(priors[: len(segments)]):
            if p is None:
                init.view(np.uint8).reshape(-1, 64)[i] = 0
    want, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, ev, off, init)
    got = I.fold(rules_of(P.bank_account_program()), 64, ev.reshape(-1, 64), off, None if init is None else init.view(np.uint8).reshape(-1, 64), f64_fields=[16])
    assert np.array_equal(got, want.view(np.uint8).reshape(got.shape))


@settings(max_examples=150, deadline=None, derandomize=True)
@given(st.lists(st.lists(st.tuples(st.integers(0, 5), st.integers(0, 70), st.integers(-2**31, 2**31 - 1)), max_size=6), min_size=1, max_size=4),
       st.integers(0, 3), st.booleans())
def test_variable_record_interpreter_equals_c_oracle(segments, cut, corrupt_len):
    """fold_var (the checker of the variable-record program fuzz) against the C oracle's VAR16 decoding of the Counter model:
    payloads of any length (too short for `by` -> the event throws), a truncated tail, a length field that lies."""
    rules = rules_of(P.counter_program(N.REC_VAR16))
    buf = bytearray()
    seg = [0]
    for i, events in enumerate(segments):
        for k, (t, plen, by) in enumerate(events):
            payload = (struct.pack("<i", by) + bytes(plen))[:plen]
            buf += struct.pack("<IIII", t, k + 1, plen, i) + payload + bytes((-plen) % 16)
        seg.append(len(buf))
    if corrupt_len and len(buf) >= 16:
        struct.pack_into("<I", buf, 8, struct.unpack_from("<I", buf, 8)[0] + 64)     # the first record claims 64 more bytes than it has
    if cut and seg[-1] - seg[-2] >= 32:
        seg[-1] -= 16                                                                  # the last segment ends inside its last record
        del buf[seg[-1]:]
    ev = np.frombuffer(bytes(buf), np.uint8) if buf else np.zeros(0, np.uint8)
    off = np.asarray(seg, dtype=np.uint64)
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_VAR16, ev, off)
    got = I.fold_var(rules, 16, ev, off)
    assert np.array_equal(got, want.view(np.uint8).reshape(got.shape))
