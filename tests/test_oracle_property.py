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


@settings(max_examples=300, deadline=None)
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
