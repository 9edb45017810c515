"""Pins the oracle against every golden vector the reference's own tests hold for the replay
path, at both levels: the object-level restatement
(oracle/surge_model.py) and the packed C restatement (oracle/sgr_oracle.c), and checks the
two restatements against each other on random small inputs.

Reference paths are relative to the reference checkout.
"""
import struct
import uuid

import numpy as np
import pytest

from oracle import oracle as O
from oracle import surge_model as M
from surge_b200 import formats as F

AGG = "cc1e6f5b-0d3c-4c26-9b5e-0c0b0d6a5a11"


# ------------------------------------------------------------------ helpers: object <-> packed
def counter_event_to_rec(evt, agg_idx=0):
    if isinstance(evt, M.CountIncremented):
        return (F.COUNT_INCREMENTED, evt.sequenceNumber, agg_idx, evt.incrementBy)
    if isinstance(evt, M.CountDecremented):
        return (F.COUNT_DECREMENTED, evt.sequenceNumber, agg_idx, evt.decrementBy)
    if isinstance(evt, M.NoOpEvent):
        return (F.NO_OP_EVENT, evt.sequenceNumber, agg_idx, 0)
    return (F.EXCEPTION_THROWING_EVENT, evt.sequenceNumber, agg_idx, 0)


def pack_counter(events, agg_idx=0):
    t = [counter_event_to_rec(e, agg_idx) for e in events]
    return F.counter_records([x[0] for x in t], [x[1] for x in t], [x[2] for x in t], [x[3] for x in t])


def counter_state_row(state, flags_extra=0):
    row = np.zeros(1, dtype=F.COUNTER_STATE)
    if state is not None:
        row["count"], row["version"], row["flags"] = state.count, state.version, O.ST_EXISTS
    row["flags"] |= flags_extra
    return row


def c_apply_counter(state, events, model=O.MODEL_COUNTER):
    rec = pack_counter(events)
    off = np.array([0, len(rec) * 64], dtype=np.uint64)
    init = counter_state_row(state)
    out, nev, nerr = O.fold_packed(model, O.REC_FIXED64, rec, off, init)
    return out.view(F.COUNTER_STATE).reshape(-1)[0], nev, nerr


# ------------------------------------------------------------------ PersistentActorSpec vectors
BASE = M.State(AGG, 3, 3)  # PersistentActorSpec.scala:181 baseState


def test_increment_from_base_state():
    """PersistentActorSpec.scala:134-168,292-308: Increment on (3,3) => CountIncremented(id,1,4) => State(id,4,4)."""
    new_state, events = M.handle_command(M.counter_process_command, M.counter_handle_event, BASE, ("Increment", AGG))
    assert events == [M.CountIncremented(AGG, 1, 4)]
    assert new_state == M.State(AGG, 4, 4)
    row, nev, nerr = c_apply_counter(BASE, events)
    assert (int(row["count"]), int(row["version"])) == (4, 4) and nev == 1 and nerr == 0
    assert int(row["flags"]) == O.ST_EXISTS | O.ST_CHANGED


def test_two_sequential_increments():
    """PersistentActorSpec.scala:466-493: (3,3) -> (4,4) -> (5,5)."""
    s = BASE
    for expect in [(4, 4), (5, 5)]:
        s, _ = M.handle_command(M.counter_process_command, M.counter_handle_event, s, ("Increment", AGG))
        assert (s.count, s.version) == expect


def test_apply_events_twice_publishes_state_twice():
    """PersistentActorSpec.scala:512-529: ApplyEvents([Incr(1,seq4)]) then ApplyEvents([Incr(1,seq5)])
    => ACK (4,4), (5,5); exactly two publishes, both state records."""
    a1 = M.apply_events(M.counter_handle_event, BASE, [M.CountIncremented(AGG, 1, 4)])
    assert a1.success and a1.state == M.State(AGG, 4, 4) and a1.published_state
    a2 = M.apply_events(M.counter_handle_event, a1.state, [M.CountIncremented(AGG, 1, 5)])
    assert a2.success and a2.state == M.State(AGG, 5, 5) and a2.published_state
    row, _, _ = c_apply_counter(BASE, [M.CountIncremented(AGG, 1, 4)])
    assert (int(row["count"]), int(row["version"]), int(row["flags"])) == (4, 4, O.ST_EXISTS | O.ST_CHANGED)
    row2, _, _ = c_apply_counter(M.State(AGG, 4, 4), [M.CountIncremented(AGG, 1, 5)])
    assert (int(row2["count"]), int(row2["version"]), int(row2["flags"])) == (5, 5, O.ST_EXISTS | O.ST_CHANGED)
    # the multi-event fold the reference never tests (PersistentActorSpec.scala:510 TODO) gives the same end state
    row3, nev, _ = c_apply_counter(BASE, [M.CountIncremented(AGG, 1, 4), M.CountIncremented(AGG, 1, 5)])
    assert (int(row3["count"]), int(row3["version"]), nev) == (5, 5, 2)


def test_unchanged_state_is_not_published():
    """PersistentActorSpec.scala:275-288: ApplyEvents([CountIncremented(id,0,3)]) on (3,3) => ACK (3,3), NO publish."""
    a = M.apply_events(M.counter_handle_event, BASE, [M.CountIncremented(AGG, 0, 3)])
    assert a.success and a.state == BASE and not a.published_state
    row, _, _ = c_apply_counter(BASE, [M.CountIncremented(AGG, 0, 3)])
    assert (int(row["count"]), int(row["version"]), int(row["flags"])) == (3, 3, O.ST_EXISTS)  # CHANGED clear


def test_noop_event_keeps_state():
    """PersistentActorSpec.scala:495-508: NoOpEvent leaves the state unchanged."""
    new_state, events = M.handle_command(M.counter_process_command, M.counter_handle_event, BASE, ("CreateNoOpEvent", AGG))
    assert events == [M.NoOpEvent(AGG, 4)] and new_state == BASE
    row, _, _ = c_apply_counter(BASE, events)
    assert (int(row["count"]), int(row["version"]), int(row["flags"])) == (3, 3, O.ST_EXISTS)


def test_noop_materialises_default_state_from_none():
    """TestBoundedContext.scala:78: agg.getOrElse(State(id,0,0)) — NoOp on None yields Some(State(id,0,0))."""
    assert M.counter_handle_event(None, M.NoOpEvent(AGG, 1)) == M.State(AGG, 0, 0)
    row, _, _ = c_apply_counter(None, [M.NoOpEvent(AGG, 1)])
    assert (int(row["count"]), int(row["version"]), int(row["flags"])) == (0, 0, O.ST_EXISTS | O.ST_CHANGED)


def test_handler_exception_keeps_previous_state():
    """PersistentActorSpec.scala:431-464: handler throws => ACKError, a later DoNothing still sees (3,3)."""
    evts = [M.CountIncremented(AGG, 7, 4), M.ExceptionThrowingEvent(AGG, 5, RuntimeError("failed"))]
    a = M.apply_events(M.counter_handle_event, BASE, evts)
    assert not a.success and a.error == "failed" and a.state == BASE and not a.published_state
    row, nev, nerr = c_apply_counter(BASE, evts)
    assert (int(row["count"]), int(row["version"])) == (3, 3)
    assert int(row["flags"]) == O.ST_EXISTS | O.ST_ERROR and int(row["err_idx"]) == 1 and nerr == 1 and nev == 1


# ------------------------------------------------------------------ multilanguage Counter
def test_multilanguage_counter_from_empty():
    """MultilanguageGatewayServiceImplSpec.scala:72-73,95-136: None -Incr-> (1,1) -Incr-> (2,2) -Decr-> (1,3)."""
    s = None
    s = M.fold_left(M.ml_counter_apply_event, s, [M.CountIncremented(AGG, 1, 1)])
    assert s == M.State(AGG, 1, 1)
    s = M.fold_left(M.ml_counter_apply_event, s, [M.CountIncremented(AGG, 1, 2)])
    assert s == M.State(AGG, 2, 2)
    s = M.fold_left(M.ml_counter_apply_event, s, [M.CountDecremented(AGG, 1, 3)])
    assert s == M.State(AGG, 1, 3)
    row, _, _ = c_apply_counter(None, [M.CountIncremented(AGG, 1, 1), M.CountIncremented(AGG, 1, 2), M.CountDecremented(AGG, 1, 3)],
                                model=O.MODEL_ML_COUNTER)
    assert (int(row["count"]), int(row["version"])) == (1, 3)
    assert M.play_json_counter_state(s) == ('{"aggregateId":"%s","count":1,"version":3}' % AGG).encode()
    assert F.counter_state_json(AGG, 1, 3) == M.play_json_counter_state(s)


def test_multilanguage_counter_rejects_other_events():
    row, _, nerr = c_apply_counter(None, [M.NoOpEvent(AGG, 1)], model=O.MODEL_ML_COUNTER)
    assert nerr == 1 and int(row["flags"]) == O.ST_ERROR


# ------------------------------------------------------------------ BankAccount
def test_bank_account_create_then_credit():
    """BankAccountCommandEngineSpec.scala:43-68: CreateAccount(n,"Jane Doe","1234",1000.0) then CreditAccount(n,100.0) => 1100.0."""
    n = str(uuid.UUID(int=0x1234))
    s, ev1 = M.handle_command(M.bank_account_process_command, M.bank_account_handle_event, None,
                              ("CreateAccount", n, "Jane Doe", "1234", 1000.0))
    assert s == M.BankAccount(n, "Jane Doe", "1234", 1000.0)
    s2, ev2 = M.handle_command(M.bank_account_process_command, M.bank_account_handle_event, s, ("CreditAccount", n, 100.0))
    assert s2 == M.BankAccount(n, "Jane Doe", "1234", 1100.0)
    # the replay path sees the two events
    assert M.fold_left(M.bank_account_handle_event, None, ev1 + ev2) == s2
    rec = F.bank_created_record(0, 1, n, "Jane Doe", "1234", 1000.0) + F.bank_updated_record(0, 2, n, 1100.0)
    out, nev, nerr = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, np.frombuffer(rec, np.uint8), np.array([0, 128], np.uint64))
    st = F.decode_bank_state(out.view(F.BANK_STATE).reshape(-1)[0])
    assert st == {"accountNumber": n, "accountOwner": "Jane Doe", "securityCode": "1234", "balance": 1100.0}
    assert (nev, nerr) == (2, 0)


def test_bank_account_update_without_account_stays_none():
    """BankAccountCommandModel.scala:84: aggregate.map(_.copy(balance = ..)) — None stays None."""
    n = str(uuid.UUID(int=7))
    assert M.bank_account_handle_event(None, M.BankAccountUpdated(n, 5.0)) is None
    rec = F.bank_updated_record(0, 1, n, 5.0)
    out, nev, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, np.frombuffer(rec, np.uint8), np.array([0, 64], np.uint64))
    assert not out.any() and nev == 1


def test_bank_account_double_equality_is_numeric():
    """Scala case-class == on Double: 0.0 == -0.0 (no publish), NaN != NaN (publish)."""
    n = str(uuid.UUID(int=9))
    base = F.bank_created_record(0, 1, n, "o", "c", 0.0)
    init, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, np.frombuffer(base, np.uint8), np.array([0, 64], np.uint64))
    upd = F.bank_updated_record(0, 2, n, -0.0)
    out, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, np.frombuffer(upd, np.uint8), np.array([0, 64], np.uint64), init)
    row = out.view(F.BANK_STATE).reshape(-1)[0]
    assert int(row["flags"]) == O.ST_EXISTS and struct.pack("<d", float(row["balance"])) == struct.pack("<d", -0.0)
    nan1 = F.bank_updated_record(0, 3, n, float("nan"))
    out2, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, np.frombuffer(nan1, np.uint8), np.array([0, 64], np.uint64), out)
    out3, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, np.frombuffer(nan1, np.uint8), np.array([0, 64], np.uint64), out2)
    assert int(out3.view(F.BANK_STATE).reshape(-1)[0]["flags"]) == O.ST_EXISTS | O.ST_CHANGED
    assert M.BankAccount(n, "o", "c", 0.0) == M.BankAccount(n, "o", "c", -0.0)


def test_untouched_instance_equals_itself_even_with_nan():
    """scalac's case-class equals is `(this eq that) || fields ==`: an ApplyEvents that hands the SAME instance back
    (no events; Counter NoOpEvent => current) publishes nothing even if a Double field holds NaN, while
    `_.copy(balance = NaN)` builds a new instance whose NaN != the old NaN (PersistentActor.scala:257)."""
    n = str(uuid.UUID(int=11))
    nan = float("nan")
    acct = M.BankAccount(n, "o", "c", nan)
    assert not M.apply_events(M.bank_account_handle_event, acct, []).published_state
    assert M.apply_events(M.bank_account_handle_event, acct, [M.BankAccountUpdated(n, nan)]).published_state
    assert M.scala_equals(acct, acct) and not M.scala_equals(acct, M.BankAccount(n, "o", "c", nan))
    st = M.State("a", 3, 3)
    assert M.counter_handle_event(st, M.NoOpEvent("a", 4)) is st
    # the C oracle: a NaN state, then an empty segment -> not CHANGED; an Updated(NaN) -> CHANGED
    created = F.bank_created_record(0, 1, n, "o", "c", nan)
    init, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, np.frombuffer(created, np.uint8), np.array([0, 64], np.uint64))
    empty, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, np.zeros(0, np.uint8), np.array([0, 0], np.uint64), init)
    assert int(empty.view(F.BANK_STATE).reshape(-1)[0]["flags"]) == O.ST_EXISTS
    upd = F.bank_updated_record(0, 2, n, nan)
    again, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, np.frombuffer(upd, np.uint8), np.array([0, 64], np.uint64), init)
    assert int(again.view(F.BANK_STATE).reshape(-1)[0]["flags"]) == O.ST_EXISTS | O.ST_CHANGED
    # the program interpreter agrees with the C oracle on both
    from oracle import program_interp as I
    rules = [(I.CREATE, [(I.OP_SET, 0, 16, 16), (I.OP_SET, 16, 32, 8), (I.OP_SET, 24, 40, 16), (I.OP_SET, 40, 56, 8)]), (I.IF_EXISTS, [(I.OP_SET, 16, 32, 8)])]
    assert np.array_equal(I.fold(rules, 64, np.zeros((0, 64), np.uint8), [0, 0], init, f64_fields=[16]), empty.view(np.uint8).reshape(1, 64))
    assert np.array_equal(I.fold(rules, 64, np.frombuffer(upd, np.uint8).reshape(1, 64), [0, 64], init, f64_fields=[16]), again.view(np.uint8).reshape(1, 64))


# ------------------------------------------------------------------ IntBalance
def test_int_balance_fold():
    """multilanguage-scala-sdk-sample Main.scala:25-30; CQRSModel.applyEvents = foldLeft(eventHandler) (scalasdk/Model.scala:9-13)."""
    s = M.fold_left(M.int_balance_event_handler, None, [M.MoneyDeposited(5), M.MoneyDeposited(2**31 - 1), M.MoneyDeposited(10)])
    assert s == M.IntBankAccount(M.jvm_int(5 + 2**31 - 1 + 10))
    rec = F.counter_records([0, 0, 0], [1, 2, 3], [0, 0, 0], [5, 2**31 - 1, 10])
    out, _, _ = O.fold_packed(O.MODEL_INT_BALANCE, O.REC_FIXED64, rec, np.array([0, 192], np.uint64))
    assert int(out.view(F.INT_BALANCE_STATE).reshape(-1)[0]["balance"]) == s.balance


# ------------------------------------------------------------------ KTable (today's recovery) + serialization wrappers
def test_ktable_last_write_wins():
    """AggregateStateStoreKafkaStreamsSpec.scala:64-85: put state1(int=1) ... then state1(int=3) => get == the latter."""
    def js(s, i):
        return ('{"string":"%s","int":%d}' % (s, i)).encode()
    t = M.ktable_restore([("state1", js("state1", 1)), ("state2", js("state2", 2)), ("state3", js("state3", 3)),
                          ("invalidValidation", js("invalidValidation", 1)), ("state1", js("state1", 3))])
    assert t["state1"] == js("state1", 3) and t["state2"] == js("state2", 2)
    assert M.ktable_restore([("a", b"x"), ("a", None)]) == {}  # null value = tombstone (SurgeModel.scala:62-64)


def test_serialized_aggregate_value_is_the_bytes():
    """SerializedAggregateSpec.scala:12-24: SerializedAggregate.create("value".getBytes).value == "value".getBytes."""
    assert bytes("value".encode()) == b"value"


# ------------------------------------------------------------------ C oracle == object-level restatement on random input
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_c_oracle_matches_object_model_counter(seed):
    rng = np.random.default_rng(seed)
    n_agg = 64
    init_states, segs = [], []
    for a in range(n_agg):
        k = int(rng.integers(0, 12))
        evs = []
        for j in range(k):
            t = rng.choice(4, p=[0.42, 0.42, 0.12, 0.04])
            by = int(rng.integers(-2**31, 2**31))
            evs.append([M.CountIncremented(AGG, by, j + 1), M.CountDecremented(AGG, by, j + 1), M.NoOpEvent(AGG, j + 1),
                        M.ExceptionThrowingEvent(AGG, j + 1, RuntimeError("x"))][t])
        segs.append(evs)
        init_states.append(None if rng.random() < 0.5 else M.State(AGG, int(rng.integers(-2**31, 2**31)), int(rng.integers(0, 100))))
    rec = np.concatenate([pack_counter(e, a) for a, e in enumerate(segs)]) if any(segs) else np.zeros(0, F.REC64)
    off = F.csr_offsets_from_counts([len(e) for e in segs])
    init = np.concatenate([counter_state_row(s) for s in init_states])
    out, nev, nerr = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, init)
    out_mt, nev2, nerr2 = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, init, threads=4)
    assert np.array_equal(out, out_mt) and (nev, nerr) == (nev2, nerr2)
    rows = out.view(F.COUNTER_STATE).reshape(-1)
    for a in range(n_agg):
        ack = M.apply_events(M.counter_handle_event, init_states[a], segs[a])
        flags = int(rows[a]["flags"])
        assert bool(flags & O.ST_ERROR) == (not ack.success)
        assert bool(flags & O.ST_CHANGED) == ack.published_state
        assert bool(flags & O.ST_EXISTS) == (ack.state is not None)
        if ack.state is not None:
            assert (int(rows[a]["count"]), int(rows[a]["version"])) == (ack.state.count, ack.state.version)
        else:
            assert int(rows[a]["count"]) == 0 and int(rows[a]["version"]) == 0


def test_partition_hash_three_restatements_agree():
    """No golden vector exists in the reference (parity unpinned): pin the three restatements to each other
    and check the documented algebra (abs of a Java remainder, takeWhile(_ != ':'))."""
    from surge_b200.partitioner import PartitionStringUpToColon, string_hash

    p = PartitionStringUpToColon()
    for s in ["", "a", "ab", "abc", "agg-17:42", AGG, AGG + ":9", "héllo wörld", "\U0001F600x", ":", "x" * 257]:
        h = M.scala_string_hash(s)
        assert h == O.scala_string_hash(s) == string_hash(s)
        for n in [1, 2, 7, 32, 1000]:
            want = M.partition_for_key(M.partition_string_up_to_colon(s), n)
            assert 0 <= want < n
            assert want == O.partition_for_key(s, n, up_to_colon=True) == p.partition_of_record_key(s, n)
    assert p.partitionBy("abc:def:ghi") == "abc"
