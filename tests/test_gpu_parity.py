"""-m gpu: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.
Bit-exact: whole state tables are compared byte for byte (integer / byte work, no tolerance).

kernel=0 lets the engine pick (fold_runs.cu for programs inside the transformer algebra),
kernel=1 forces the lane-sequential TMA kernel (fold_kernels.cu), kernel=3 the record-per-lane
rows kernel (fold_rows.cu).
"""
import uuid

import numpy as np
import pytest

from oracle import oracle as O
from oracle import surge_model as M
from surge_b200 import ReplayEngine, SgrError
from surge_b200 import formats as F
from surge_b200 import native as N
from surge_b200 import programs as P
from surge_b200 import synth as S

pytestmark = pytest.mark.gpu

KERNELS = [0, 1, 3]


def run_engine(prog, events, offsets, kernel=0, init=None, variant=None, run_variant=None):
    with ReplayEngine(0) as e:
        e.register_program(prog)
        e.set_option("kernel", kernel)
        if variant is not None:
            e.set_option("fold_variant", variant)
        if run_variant is not None:
            e.set_option("run_variant", run_variant)
        if init is not None:
            e.set_initial_states(init)
        e.load_events(events, offsets)
        e.fold()
        st = e.stats()
        return e.export_states(), st


def assert_same(got, want, what=""):
    if not np.array_equal(got, want):
        bad = np.nonzero((got != want).any(axis=1))[0]
        raise AssertionError(f"{what}: {len(bad)} of {len(want)} states differ; first {bad[:8]}\n"
                             f"got  {got[bad[:4]].tolist()}\nwant {want[bad[:4]].tolist()}")


# ------------------------------------------------------------------ Counter, fixed records
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("n_agg,max_events,seed", [(1, 1, 1), (7, 3, 2), (1000, 40, 3), (20000, 70, 4), (300, 600, 5)])
def test_counter_ragged_segments(kernel, n_agg, max_events, seed):
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, max_events + 1, size=n_agg)
    rec, off = S.counter_csr(n_agg, counts, seed=seed, p_throw=0.003)
    want, nev, nerr = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    got, st = run_engine(P.counter_program(), rec, off, kernel)
    assert_same(got, want, f"kernel {kernel}")
    assert (st.n_events, st.n_errors) == (nev, nerr)


@pytest.mark.parametrize("kernel", KERNELS)
def test_counter_with_prior_states_and_publish_rule(kernel):
    rng = np.random.default_rng(11)
    n_agg = 5000
    counts = rng.integers(0, 20, size=n_agg)
    rec, off = S.counter_csr(n_agg, counts, seed=12, p_throw=0.01, by_max=3)  # small `by`: unchanged states do occur
    init = np.zeros(n_agg, dtype=F.COUNTER_STATE)
    ex = rng.random(n_agg) < 0.6
    init["count"][ex] = rng.integers(-5, 5, size=ex.sum())
    init["version"][ex] = rng.integers(0, 20, size=ex.sum())
    init["flags"][ex] = N.ST_EXISTS
    want, nev, nerr = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, init)
    got, st = run_engine(P.counter_program(), rec, off, kernel, init=init)
    assert_same(got, want)
    flags = want.view(F.COUNTER_STATE).reshape(-1)["flags"]
    assert ((flags & N.ST_CHANGED) == 0).any() and (flags & N.ST_CHANGED).any() and (flags & N.ST_ERROR).any()


@pytest.mark.parametrize("kernel", KERNELS)
def test_counter_one_long_segment_and_neighbours(kernel):
    """A hot aggregate much longer than a warp span, between short ones (skew): exercises the
    cross-span look-back of the record-parallel kernel and the chunk ring of the sequential one."""
    counts = np.array([3, 0, 200_000, 1, 0, 0, 5, 40_000, 2])
    rec, off = S.counter_csr(len(counts), counts, seed=21)
    want, nev, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    got, st = run_engine(P.counter_program(), rec, off, kernel)
    assert_same(got, want)
    assert st.n_events == nev


@pytest.mark.parametrize("kernel", KERNELS)
def test_counter_long_segment_with_late_throw(kernel):
    counts = np.array([10, 150_000, 10])
    rec, off = S.counter_csr(3, counts, seed=22)
    rec["type"][10 + 149_990] = F.EXCEPTION_THROWING_EVENT
    want, nev, nerr = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    got, st = run_engine(P.counter_program(), rec, off, kernel)
    assert_same(got, want)
    assert nerr == 1 and st.n_errors == 1
    row = got.view(F.COUNTER_STATE).reshape(-1)[1]
    assert int(row["flags"]) == N.ST_ERROR and int(row["err_idx"]) == 149_990


@pytest.mark.parametrize("kernel", KERNELS)
def test_counter_uniform_config2_shape_small(kernel):
    """configs[1] shape at a size the oracle folds in a second: 16384 aggregates x 32 events."""
    rec, off = S.counter_csr(16384, 32, seed=2)
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, threads=4)
    got, _ = run_engine(P.counter_program(), rec, off, kernel)
    assert_same(got, want)


@pytest.mark.parametrize("kernel", KERNELS)
def test_log_not_starting_at_zero_and_all_empty(kernel):
    rec, off = S.counter_csr(50, 4, seed=31)
    # the CSR may address a sub-range of a larger buffer
    pad = np.zeros(3, dtype=F.REC64)
    buf = np.concatenate([pad, rec, pad])
    off2 = off + np.uint64(3 * 64)
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, buf, off2)
    got, _ = run_engine(P.counter_program(), buf, off2, kernel)
    assert_same(got, want)
    # all segments empty: every state is None
    offe = np.zeros(65, dtype=np.uint64)
    got, st = run_engine(P.counter_program(), np.zeros(1, dtype=F.REC64), offe, kernel)
    assert not got.any() and st.n_events == 0


@pytest.mark.parametrize("variant", range(6))
def test_sequential_kernel_variants(variant):
    rng = np.random.default_rng(40 + variant)
    counts = rng.integers(0, 50, size=3000)
    rec, off = S.counter_csr(len(counts), counts, seed=41, p_throw=0.002)
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    got, _ = run_engine(P.counter_program(), rec, off, kernel=1, variant=variant)
    assert_same(got, want)


@pytest.mark.parametrize("run_variant", range(7))
def test_runs_kernel_variants(run_variant):
    rng = np.random.default_rng(50 + run_variant)
    counts = np.concatenate([rng.integers(0, 50, size=3000), [60_000], rng.integers(0, 4, size=500)])
    rec, off = S.counter_csr(len(counts), counts, seed=51, p_throw=0.001)
    init = np.zeros(len(counts), dtype=F.COUNTER_STATE)
    ex = rng.random(len(counts)) < 0.5
    init["count"][ex] = 7
    init["flags"][ex] = N.ST_EXISTS
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, init)
    got, _ = run_engine(P.counter_program(), rec, off, kernel=2, init=init, run_variant=run_variant)
    assert_same(got, want)


# ------------------------------------------------------------------ golden vectors through the GPU path
def _counter_row(count, version, flags):
    r = np.zeros(1, dtype=F.COUNTER_STATE)
    r["count"], r["version"], r["flags"] = count, version, flags
    return r


@pytest.mark.parametrize("kernel", KERNELS)
def test_reference_golden_vectors_on_gpu(kernel):
    """SURVEY.md Appendix D vectors, each as one aggregate of one batch:
    PersistentActorSpec.scala:134-168,275-288,431-464,466-529; MultilanguageGatewayServiceImplSpec.scala:72-136."""
    base = (3, 3, N.ST_EXISTS)
    cases = [  # (prior, [(type, seq, by)], expected (count, version, flags, err_idx))
        (base, [(0, 4, 1)], (4, 4, N.ST_EXISTS | N.ST_CHANGED, 0)),
        (base, [(0, 4, 1), (0, 5, 1)], (5, 5, N.ST_EXISTS | N.ST_CHANGED, 0)),
        (base, [(0, 3, 0)], (3, 3, N.ST_EXISTS, 0)),                       # unchanged => no publish
        (base, [(2, 4, 0)], (3, 3, N.ST_EXISTS, 0)),                       # NoOpEvent
        (base, [(0, 4, 7), (3, 5, 0)], (3, 3, N.ST_EXISTS | N.ST_ERROR, 1)),  # handler throws => state kept
        (None, [(0, 1, 1), (0, 2, 1), (1, 3, 1)], (1, 3, N.ST_EXISTS | N.ST_CHANGED, 0)),  # multilanguage counter
        (None, [(2, 1, 0)], (0, 0, N.ST_EXISTS | N.ST_CHANGED, 0)),       # NoOp materialises State(id,0,0)
        (None, [], (0, 0, 0, 0)),
    ]
    init = np.concatenate([_counter_row(*(c[0] or (0, 0, 0))) for c in cases])
    recs = [F.counter_records([t for t, _, _ in ev], [s for _, s, _ in ev], [i] * len(ev), [b for _, _, b in ev]) for i, (_, ev, _) in enumerate(cases)]
    rec = np.concatenate(recs)
    off = F.csr_offsets_from_counts([len(c[1]) for c in cases])
    got, _ = run_engine(P.counter_program(), rec, off, kernel, init=init)
    rows = got.view(F.COUNTER_STATE).reshape(-1)
    for i, (_, _, exp) in enumerate(cases):
        assert (int(rows[i]["count"]), int(rows[i]["version"]), int(rows[i]["flags"]), int(rows[i]["err_idx"])) == exp, i
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, init)
    assert_same(got, want)


# ------------------------------------------------------------------ other models
@pytest.mark.parametrize("kernel", [2, 1])
def test_bank_account_model(kernel):
    """BankAccount (f64 copy, strings, IF_EXISTS rule): BankAccountCommandEngineSpec.scala:43-68 + random logs.
    kernel 2: the 64-byte class-1 instantiation of fold_runs.cu; kernel 1: the lane-sequential kernel
    (kernel 0 would pick the latter for a balanced log of 64-byte states)."""
    rng = np.random.default_rng(5)
    n_agg = 2000
    blobs, counts = [], []
    for a in range(n_agg):
        acct = str(uuid.UUID(int=int(rng.integers(1, 2**62))))
        k = int(rng.integers(0, 12))
        evs = []
        for j in range(k):
            if rng.random() < 0.25:
                evs.append(F.bank_created_record(a, j + 1, acct, f"owner{a % 97}", f"{a % 10000:04d}", 1000.0 + 0.25 * j))
            else:
                bal = [float(j), -0.0, 0.0, float("nan"), 1e300][int(rng.integers(0, 5))]
                evs.append(F.bank_updated_record(a, j + 1, acct, bal))
        if a == 0:
            n0 = str(uuid.UUID(int=0x1234))
            evs = [F.bank_created_record(0, 1, n0, "Jane Doe", "1234", 1000.0), F.bank_updated_record(0, 2, n0, 1100.0)]
        blobs.append(b"".join(evs))
        counts.append(len(evs))
    rec = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    off = F.csr_offsets_from_counts(counts)
    want, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, rec, off)
    got, st = run_engine(P.bank_account_program(), rec, off, kernel)
    assert_same(got, want)
    assert st.fold_launches == 1
    assert F.decode_bank_state(got.view(F.BANK_STATE).reshape(-1)[0]) == {
        "accountNumber": str(uuid.UUID(int=0x1234)), "accountOwner": "Jane Doe", "securityCode": "1234", "balance": 1100.0}
    # second batch on top: the publish rule with JVM Double equality (0.0 == -0.0, NaN != NaN)
    rec2, cnt2 = [], []
    for a in range(n_agg):
        acct = str(uuid.UUID(int=a + 1))
        bal = [0.0, -0.0, float("nan"), 5.0][a % 4]
        rec2.append(F.bank_updated_record(a, 100, acct, bal))
        cnt2.append(1)
    rec2 = np.frombuffer(b"".join(rec2), dtype=np.uint8)
    off2 = F.csr_offsets_from_counts(cnt2)
    want2, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, rec2, off2, want)
    got2, _ = run_engine(P.bank_account_program(), rec2, off2, kernel, init=got)
    assert_same(got2, want2)
    want3, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, rec2, off2, want2)
    got3, _ = run_engine(P.bank_account_program(), rec2, off2, kernel, init=got2)
    assert_same(got3, want3)


@pytest.mark.parametrize("kernel", KERNELS)
def test_int_balance_and_ml_counter(kernel):
    rng = np.random.default_rng(6)
    counts = rng.integers(0, 30, size=4000)
    rec, off = S.counter_csr(len(counts), counts, seed=61)
    rec_ib = rec.copy()
    rec_ib["type"] = np.where(rng.random(len(rec)) < 0.002, 1, 0)  # type 1 is a MatchError for IntBalance
    want, _, _ = O.fold_packed(O.MODEL_INT_BALANCE, O.REC_FIXED64, rec_ib, off)
    got, _ = run_engine(P.int_balance_program(), rec_ib, off, kernel)
    assert_same(got, want)
    want, _, nerr = O.fold_packed(O.MODEL_ML_COUNTER, O.REC_FIXED64, rec, off)  # NoOp (type 2) is a MatchError here
    got, st = run_engine(P.ml_counter_program(), rec, off, kernel)
    assert_same(got, want)
    assert nerr > 0 and st.n_errors == nerr


# ------------------------------------------------------------------ variable records (config 4 shape)
def test_counter_variable_records():
    rng = np.random.default_rng(7)
    counts = rng.integers(0, 25, size=3000)
    buf, seg = S.counter_var_csr(len(counts), counts, seed=71)
    want, nev, nerr = O.fold_packed(O.MODEL_COUNTER, O.REC_VAR16, buf, seg)
    got, st = run_engine(P.counter_program(N.REC_VAR16), buf, seg)
    assert_same(got, want)
    assert (st.n_events, st.n_errors) == (nev, nerr)


@pytest.mark.parametrize("skew", [False, True])
def test_counter_variable_records_with_directory(skew):
    """configs[3] shape, small: variable records + record directory -> record-parallel kernel (fold_vruns.cu).
    skew=True puts ~40 % of all events on one hot aggregate (crosses many warp spans) and adds throwing events."""
    rng = np.random.default_rng(8)
    counts = rng.integers(0, 25, size=4000)
    if skew:
        counts[1234] = 60_000
    buf, seg, rec_off = S.counter_var_csr(len(counts), counts, seed=72, with_directory=True)
    if skew:  # a few throwing events (type 3), one inside the hot aggregate
        for j in [5, 1000, int(np.searchsorted(rec_off, seg[1234])) + 59_000]:
            buf[int(rec_off[j]):int(rec_off[j]) + 4] = np.frombuffer(np.uint32(3).tobytes(), np.uint8)
    want, nev, nerr = O.fold_packed(O.MODEL_COUNTER, O.REC_VAR16, buf, seg)
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program(N.REC_VAR16))
        e.load_events_indexed(buf, seg, rec_off)
        e.fold()
        got, st = e.export_states(), e.stats()
    assert_same(got, want)
    assert (st.n_events, st.n_errors) == (nev, nerr) and st.fold_launches == 2   # vruns + replay, not the lane-per-aggregate kernel


def test_variable_records_directory_disagreeing_with_csr_falls_back():
    """The CSR is the source of truth: if header aggregate indices do not match it, the engine folds sequentially."""
    counts = np.array([3, 4, 5, 2])
    buf, seg, rec_off = S.counter_var_csr(4, counts, seed=73, with_directory=True)
    buf[int(rec_off[4]) + 12: int(rec_off[4]) + 16] = np.frombuffer(np.uint32(0).tobytes(), np.uint8)  # record 4 claims aggregate 0 instead of 1
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_VAR16, buf, seg)
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program(N.REC_VAR16))
        e.load_events_indexed(buf, seg, rec_off)
        e.fold()
        assert_same(e.export_states(), want)


def test_variable_records_malformed():
    """Too-short payload for the event class, a record running past its segment, and an over-long record."""
    segs = []
    t, s, a, p = [0, 0], [1, 2], [0, 0], [b"\x05\0\0\0" + bytes(28), b"\x01\0"]       # second record: payload < 4 bytes
    segs.append(F.pack_var_records(t, s, a, p)[0])
    b1 = F.pack_var_records([0], [1], [1], [b"\x07\0\0\0" + bytes(60)])[0].copy()
    b1[8:12] = np.frombuffer(np.uint32(4000).tobytes(), np.uint8)                       # claims more than the segment holds
    segs.append(b1)
    segs.append(F.pack_var_records([1, 2], [1, 2], [2, 2], [b"\x03\0\0\0" + bytes(44), bytes(32)])[0])  # fine
    buf = np.concatenate(segs)
    seg = np.zeros(4, dtype=np.uint64)
    np.cumsum([len(x) for x in segs], out=seg[1:])
    want, _, nerr = O.fold_packed(O.MODEL_COUNTER, O.REC_VAR16, buf, seg)
    got, st = run_engine(P.counter_program(N.REC_VAR16), buf, seg)
    assert_same(got, want)
    assert nerr == 2 == st.n_errors


# ------------------------------------------------------------------ K5 group-by and K6 incremental
@pytest.mark.parametrize("n_agg,max_events,seed", [(10, 5, 1), (5000, 30, 2), (70000, 6, 3)])
def test_unsorted_load_groups_stably(n_agg, max_events, seed):
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, max_events + 1, size=n_agg)
    rec, off = S.counter_csr(n_agg, counts, seed=seed + 100)
    arrival = S.interleave_arrival(rec, seed=seed + 200)
    grouped, goff = O.group_by_agg(arrival, n_agg)
    assert np.array_equal(grouped.reshape(-1), rec.view(np.uint8).reshape(-1)) and np.array_equal(goff, off)  # oracle group-by == CSR
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        e.load_unsorted(arrival, n_agg)
        e.fold()
        assert_same(e.export_states(), want)
        assert e.stats().n_events == len(rec)


def test_unsorted_load_rejects_out_of_range_aggregate():
    rec, _ = S.counter_csr(10, 3, seed=1)
    rec["agg"][5] = 99
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        with pytest.raises(SgrError) as ei:
            e.load_unsorted(rec, 10)
        assert ei.value.code == N.SGR_ERR_INVALID


@pytest.mark.parametrize("path", [0, 1])
def test_incremental_micro_batches(path):
    """configs[4] shape, small: batches appended to live aggregates; each batch == one ApplyEvents per touched aggregate.
    path 0: sort-free atomic K6 (incremental.cu); path 1: sort-based K5 + fold."""
    n_agg = 20000
    rec, off = S.counter_csr(n_agg, 3, seed=81)
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    rng = np.random.default_rng(82)
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        e.set_option("incremental", path)
        e.load_events(rec, off)
        e.fold()
        for b in range(8):
            n = [1000, 1, 5000, 0, 300, 20000, 64, 3000][b]
            aggs = rng.integers(0, n_agg, size=n).astype(np.uint64)
            types = rng.choice([0, 1, 2, 3], size=n, p=[0.45, 0.44, 0.1, 0.01]).astype(np.uint32)
            if b == 6:
                aggs[:] = 17          # one hot aggregate, with a throw in the middle: err_idx must be exact
                types[:] = 0
                types[40] = 3
            batch = F.counter_records(types, np.arange(n, dtype=np.uint32) + 1000 * b, aggs, rng.integers(0, 3, size=n).astype(np.int32))
            want = O.fold_incremental(O.MODEL_COUNTER, batch, want)
            e.fold_incremental(batch)
            assert_same(e.export_states(), want, f"batch {b}")
            st = e.stats()
            errs = int((want.view(F.COUNTER_STATE).reshape(-1)["flags"] & N.ST_ERROR != 0).sum())
            assert st.n_errors == errs and st.n_aggregates == len(np.unique(aggs))
        bad = F.counter_records([0], [1], [n_agg + 5], [1])
        with pytest.raises(SgrError):
            e.fold_incremental(bad)
        assert_same(e.export_states(), O.fold_incremental(O.MODEL_COUNTER, np.zeros(0, F.REC64), want), "rejected batch leaves the table alone")


# ------------------------------------------------------------------ recovery read (getAggregateBytes)
def test_get_aggregate_bytes_by_key():
    n_agg = 500
    rec, off = S.counter_csr(n_agg, 5, seed=91)
    rec["type"][0:5] = F.EXCEPTION_THROWING_EVENT  # aggregate 0 stays None (and in error)
    keys = [f"agg-{i:05d}" for i in range(n_agg)]
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        e.load_keys(keys)
        with pytest.raises(N.InvalidStateStoreException):
            e.get(keys[3])  # nothing folded yet: the store is not readable
        e.load_events(rec, off)
        e.fold()
        assert e.get(keys[0]) is None
        assert e.get("no-such-aggregate") is None
        for i in [1, 2, 250, 499]:
            assert e.get(keys[i]) == bytes(want[i][:8])
        b, flags, err = e.get_index(0)
        assert b is None and flags == N.ST_ERROR and err == 0


# ------------------------------------------------------------------ full config-2 size: size-independent properties
def test_config2_full_size_properties():
    """1,048,576 aggregates x 32 x 64-byte events (configs[1]). The oracle would take minutes in Python-driven
    pieces, so: (1) the whole table against a vectorised torch restatement of the Counter algebra
    (count = sum of +-by mod 2^32, version = last non-NoOp seq); (2) a random sample of aggregates against the oracle;
    (3) idempotence: folding the same log again from None gives identical bytes; both kernels agree."""
    import torch

    n_agg, epa = 1 << 20, 32
    rec, off = S.counter_csr_device(n_agg, epa, seed=2)
    r = rec.view(n_agg, epa, 16)
    t, by = r[:, :, 0], r[:, :, 4].to(torch.int64)
    cnt = torch.where(t == 0, by, torch.where(t == 1, -by, torch.zeros_like(by))).sum(1)
    cnt = ((cnt + (1 << 31)) % (1 << 32) - (1 << 31)).to(torch.int32)
    ver = torch.where(t != 2, r[:, :, 1], torch.zeros_like(t)).max(1).values
    tables = []
    for kernel in KERNELS:
        with ReplayEngine(0) as e:
            e.register_program(P.counter_program())
            e.set_option("kernel", kernel)
            e.load_events(rec.view(torch.uint8), off)
            e.fold()
            st = e.states_tensor().view(torch.int32).view(n_agg, 4)
            assert bool((st[:, 0] == cnt).all()) and bool((st[:, 1] == ver).all())
            assert bool((st[:, 2] == (N.ST_EXISTS | N.ST_CHANGED)).all()) and bool((st[:, 3] == 0).all())
            assert e.stats().n_events == n_agg * epa
            first = e.export_states()
            e.set_initial_states(None)
            e.fold()
            assert np.array_equal(first, e.export_states())
            tables.append(first)
    assert np.array_equal(tables[0], tables[1])
    sample = np.random.default_rng(3).choice(n_agg, size=4096, replace=False)
    sample.sort()
    host = rec.view(n_agg, epa * 16)[torch.as_tensor(sample, device=rec.device)].cpu().numpy().view(np.uint8).reshape(-1)
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, host, F.csr_offsets_from_counts([epa] * len(sample)))
    assert np.array_equal(tables[0][sample], want)


# ------------------------------------------------------------------ a7: KTable restore of the state topic (last write wins)
@pytest.mark.parametrize("kernel", [0, 1])
def test_state_topic_restore_is_last_write_wins(kernel):
    """AggregateStateStoreKafkaStreamsSpec.scala:64-85 at scale: replaying state snapshots in arrival order leaves, per key,
    the bytes of the LAST snapshot; a null value (tombstone) deletes the key. Checked against the object-level KTable restatement."""
    rng = np.random.default_rng(77)
    n_agg, n = 3000, 40000
    aggs = rng.integers(0, n_agg, size=n)
    tomb = rng.random(n) < 0.08
    counts = rng.integers(-2**31, 2**31, size=n)
    versions = rng.integers(0, 1000, size=n)
    rec = F.counter_records(tomb.astype(np.uint32), np.arange(n, dtype=np.uint32), aggs.astype(np.uint64), counts.astype(np.int32))
    rec["arg1"] = versions.astype(np.int32)
    table = M.ktable_restore([(str(a), None if t else (int(c), int(v))) for a, t, c, v in zip(aggs, tomb, counts, versions)])
    with ReplayEngine(0) as e:
        e.register_program(P.counter_snapshot_restore_program())
        e.set_option("kernel", kernel)
        e.load_unsorted(rec, n_agg)
        e.fold()
        rows = e.export_states().view(F.COUNTER_STATE).reshape(-1)
    for a in range(n_agg):
        want = table.get(str(a))
        if want is None:
            assert int(rows[a]["flags"]) & N.ST_EXISTS == 0 and int(rows[a]["count"]) == 0
        else:
            assert (int(rows[a]["count"]), int(rows[a]["version"])) == want and int(rows[a]["flags"]) & N.ST_EXISTS


# ------------------------------------------------------------------ API behaviour
def test_pipelined_async_folds_and_wait():
    rec, off = S.counter_csr(5000, 16, seed=111)
    want, nev, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        e.load_events(rec, off)
        for _ in range(5):
            e.set_initial_states(None)   # every fold is a rebuild from None; no host round trip in between
            e.fold_async()
        e.wait()
        assert_same(e.export_states(), want)
        assert e.stats().n_events == nev
        # without the reset a second fold appends the same log onto the live table (ApplyEvents on live actors)
        e.fold()
        want2, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off, want)
        assert_same(e.export_states(), want2)


def test_error_statuses():
    with ReplayEngine(0) as e:
        with pytest.raises(SgrError) as ei:
            e.load_events(np.zeros(64, np.uint8), np.array([0, 64], np.uint64))
        assert ei.value.code == N.SGR_ERR_NO_PROGRAM
        bad = P.counter_program()
        bad.state_bytes = 20
        with pytest.raises(SgrError) as ei:
            e.register_program(bad)
        assert ei.value.code == N.SGR_ERR_INVALID
        bad = P.counter_program()
        bad.rules[0].ops[0].src_off = 62   # reads past the 64-byte record
        with pytest.raises(SgrError) as ei:
            e.register_program(bad)
        assert ei.value.code == N.SGR_ERR_INVALID
        e.register_program(P.counter_program())
        with pytest.raises(SgrError) as ei:
            e.fold()
        assert ei.value.code == N.SGR_ERR_NOT_LOADED
        with pytest.raises(SgrError) as ei:
            e.load_events(np.zeros(128, np.uint8), np.array([0, 24, 128], np.uint64))   # offsets must be multiples of 16
        assert ei.value.code == N.SGR_ERR_INVALID
        with pytest.raises(SgrError) as ei:
            e.fold_incremental(np.zeros(64, np.uint8))   # no live table yet
        assert ei.value.code == N.SGR_ERR_NOT_LOADED
        rec, off = S.counter_csr(10, 2, seed=1)
        e.load_events(rec, off)
        e.fold()
        small = np.zeros((5, 16), np.uint8)
        with pytest.raises(SgrError) as ei:
            e.export_states(small)
        assert ei.value.code == N.SGR_ERR_CAPACITY
        e.load_keys([f"k{i}" for i in range(10)])
        with pytest.raises(SgrError):
            e.load_keys(["dup", "dup"])


def test_bank_account_incremental_takes_the_sort_based_path():
    """A program outside the transformer algebra (IF_EXISTS, 64-byte state) appended as micro-batches."""
    n_agg = 300
    accts = [str(uuid.UUID(int=i + 1)) for i in range(n_agg)]
    rng = np.random.default_rng(9)
    first = b"".join(F.bank_created_record(a, 1, accts[a], f"o{a}", "1", 10.0) for a in range(0, n_agg, 2))   # even accounts exist
    with ReplayEngine(0) as e:
        e.register_program(P.bank_account_program())
        e.set_initial_states(np.zeros((n_agg, 64), np.uint8))
        e.fold_incremental(np.frombuffer(first, np.uint8))
        want = O.fold_incremental(O.MODEL_BANK_ACCOUNT, np.frombuffer(first, np.uint8), np.zeros((n_agg, 64), np.uint8))
        assert_same(e.export_states(), want)
        for b in range(3):
            aggs = rng.integers(0, n_agg, size=500)
            batch = b"".join(F.bank_updated_record(int(a), 10 + b, accts[int(a)], float(b) + 0.5 * int(a)) for a in aggs)
            want = O.fold_incremental(O.MODEL_BANK_ACCOUNT, np.frombuffer(batch, np.uint8), want)
            e.fold_incremental(np.frombuffer(batch, np.uint8))
            assert_same(e.export_states(), want, f"batch {b}")
        rows = want.view(F.BANK_STATE).reshape(-1)
        assert not (rows["flags"][1::2] & N.ST_EXISTS).any()   # updates never create an account


def test_single_rank_route_and_fold():
    """The multi-GPU entry point on one rank: ownership tables from the real partitioner, exchange degenerates to the group-by."""
    from surge_b200 import dist as D

    n_global = 3000
    counts = np.random.default_rng(3).integers(0, 9, size=n_global)
    rec, off = S.counter_csr(n_global, counts, seed=31, p_throw=0.002)
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    arrival = S.interleave_arrival(rec, seed=32)
    import torch

    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        e.dist_init(0, 1, None, len(arrival) + 10)
        e.dist_set_partitions(D.partitions_for_keys([f"agg-{g}" for g in range(n_global)], 32))
        e.dist_route_and_fold(torch.from_numpy(arrival.view(np.uint8).reshape(-1).copy()).cuda(), fused=False)
        gl = e.dist_local_aggregates()
        assert np.array_equal(gl, np.arange(n_global, dtype=np.uint32))
        assert_same(e.export_states(), want)
        assert e.dist_stats().n_recv == len(arrival)


def test_bank_account_long_segments_cross_spans():
    """IF_EXISTS composition across lanes, steps and warp spans: accounts with thousands of updates, some before the
    account exists, creations in the middle, and unknown event types (MatchError) in a long segment."""
    rng = np.random.default_rng(15)
    counts = [3, 30_000, 0, 12_000, 7, 25_000]
    blobs = []
    for a, k in enumerate(counts):
        acct = str(uuid.UUID(int=a + 1))
        evs = []
        created_at = {1: 0, 3: 5000, 5: None}.get(a, 0)      # account 3 is created late, account 5 never
        for j in range(k):
            if created_at is not None and j == created_at:
                evs.append(F.bank_created_record(a, j + 1, acct, f"owner{a}", "42", 100.0))
            else:
                evs.append(F.bank_updated_record(a, j + 1, acct, float(j) * 0.5))
        if a == 1:
            bad = bytearray(evs[29_000]); bad[0:4] = (7).to_bytes(4, "little"); evs[29_000] = bytes(bad)   # MatchError late in the segment
        blobs.append(b"".join(evs))
    rec = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    off = F.csr_offsets_from_counts(counts)
    want, _, nerr = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, rec, off)
    for kernel in (0, 1):
        got, st = run_engine(P.bank_account_program(), rec, off, kernel)
        assert_same(got, want, f"kernel {kernel}")
        assert st.n_errors == nerr == 1
    rows = want.view(F.BANK_STATE).reshape(-1)
    assert int(rows[1]["flags"]) == N.ST_ERROR and int(rows[1]["err_idx"]) == 29_000
    assert not int(rows[5]["flags"]) & N.ST_EXISTS and float(rows[3]["balance"]) == (12_000 - 1) * 0.5


def _wide_counter_program():
    """A 32-byte state outside the sample models: count/version as Counter, plus last `by` (SET), running sum of seq (ADD),
    and two payload words copied on every counting event; NoOp materialises; type 3 resets everything (CREATE)."""
    return P.make_program(32, N.REC_FIXED64, [
        (N.MATERIALISE, [(N.OP_ADD_I32, 0, 16, 4), (N.OP_SET, 4, 4, 4), (N.OP_SET, 8, 16, 4), (N.OP_ADD_I32, 12, 4, 4), (N.OP_SET, 16, 32, 8)]),
        (N.MATERIALISE, [(N.OP_SUB_I32, 0, 16, 4), (N.OP_SET, 4, 4, 4), (N.OP_SET, 8, 16, 4), (N.OP_ADD_I32, 12, 4, 4)]),
        (N.MATERIALISE, []),
        (N.CREATE, [(N.OP_SET, 0, 16, 4)]),
        (N.TOMBSTONE, []),
    ])


def test_wide_state_program_matches_the_sequential_kernel():
    """32-byte states on fold_runs.cu (W = 6): no oracle model has this shape, so the lane-sequential kernel — itself
    pinned to the oracle on every other program — is the reference."""
    rng = np.random.default_rng(16)
    counts = np.concatenate([rng.integers(0, 60, size=3000), [40_000]])
    rec, off = S.counter_csr(len(counts), counts, seed=161)
    rec["type"] = rng.choice([0, 1, 2, 3, 4, 9], size=len(rec), p=[0.4, 0.4, 0.1, 0.05, 0.04, 0.01]).astype(np.uint32)
    rec["pad"][:, :8] = rng.integers(0, 256, size=(len(rec), 8), dtype=np.uint8)
    got0, st0 = run_engine(_wide_counter_program(), rec, off, kernel=0)
    got1, st1 = run_engine(_wide_counter_program(), rec, off, kernel=1)
    assert_same(got0, got1)
    assert (st0.n_events, st0.n_errors) == (st1.n_events, st1.n_errors) and st0.n_errors > 0
    init = got1.copy()
    got0b, _ = run_engine(_wide_counter_program(), rec, off, kernel=0, init=init)
    got1b, _ = run_engine(_wide_counter_program(), rec, off, kernel=1, init=init)
    assert_same(got0b, got1b)


def test_micro_batch_with_many_throwing_slots_takes_the_deferred_replay():
    """K6: when re-scanning the batch per throwing slot would cost too much, the throwing slots are replayed through the
    sort-based path instead (forced here with a zero budget)."""
    n_agg = 5000
    rec, off = S.counter_csr(n_agg, 2, seed=171)
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    rng = np.random.default_rng(172)
    n = 30000
    batch = F.counter_records(rng.choice([0, 1, 2, 3], size=n, p=[0.4, 0.4, 0.1, 0.1]).astype(np.uint32), np.arange(n, dtype=np.uint32),
                              rng.integers(0, n_agg, size=n).astype(np.uint64), rng.integers(0, 9, size=n).astype(np.int32))
    want = O.fold_incremental(O.MODEL_COUNTER, batch, want)
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        e.set_option("replay_budget", 0)
        e.load_events(rec, off)
        e.fold()
        e.fold_incremental(batch)
        assert_same(e.export_states(), want)
        st = e.stats()
        nerr = int((want.view(F.COUNTER_STATE).reshape(-1)["flags"] & N.ST_ERROR != 0).sum())
        assert st.n_errors == nerr > 1000


def test_micro_batches_with_a_word_that_is_both_set_and_added():
    """K6 general (three-phase) mode: a state word that sees SETs and ADDs needs the 'adds after the last SET' rule.
    No oracle model has this shape; the sort-based path (K5 + sequential kernel, pinned elsewhere) is the reference."""
    prog = P.make_program(16, N.REC_FIXED64, [
        (N.MATERIALISE, [(N.OP_ADD_I32, 0, 16, 4), (N.OP_SET, 4, 4, 4)]),      # count += by, version = seq
        (N.MATERIALISE, [(N.OP_SET, 0, 16, 4)]),                               # count = by
        (N.TOMBSTONE, []),
        (N.CREATE, [(N.OP_SUB_I32, 0, 16, 4)]),                                # reset, count = -by
    ])
    n_agg = 3000
    rng = np.random.default_rng(181)
    tables = []
    for path in (0, 1):
        rng = np.random.default_rng(181)
        with ReplayEngine(0) as e:
            e.register_program(prog)
            e.set_option("incremental", path)
            e.set_initial_states(np.zeros((n_agg, 16), np.uint8))
            for b in range(4):
                n = 20000
                batch = F.counter_records(rng.choice([0, 1, 2, 3, 7], size=n, p=[0.6, 0.2, 0.05, 0.1, 0.05]).astype(np.uint32),
                                          np.arange(n, dtype=np.uint32) + 7 * b, rng.integers(0, n_agg, size=n).astype(np.uint64),
                                          rng.integers(-50, 50, size=n).astype(np.int32))
                e.fold_incremental(batch)
            tables.append(e.export_states())
    assert_same(tables[0], tables[1])
    assert (tables[0].view(F.COUNTER_STATE).reshape(-1)["flags"] & N.ST_ERROR).any()


@pytest.mark.parametrize("prog_name", ["counter", "bank"])
def test_fold_unsorted_one_call(prog_name):
    """sgr_fold_unsorted: an arrival-order log straight to states. Counter takes the sort-free atomic fold, BankAccount the
    group-by + fold; both must equal the oracle's fold of the grouped log, and a following micro-batch must append correctly."""
    rng = np.random.default_rng(191)
    n_agg = 4000
    if prog_name == "counter":
        counts = rng.integers(0, 30, size=n_agg)
        rec, off = S.counter_csr(n_agg, counts, seed=192, p_throw=0.003)
        model, prog = O.MODEL_COUNTER, P.counter_program()
        arrival = S.interleave_arrival(rec, seed=193)
    else:
        blobs, counts = [], []
        for a in range(n_agg):
            acct = str(uuid.UUID(int=a + 1)); k = int(rng.integers(0, 6))
            evs = [F.bank_created_record(a, 1, acct, "o", "c", 1.0)] if k and rng.random() < 0.7 else []
            evs += [F.bank_updated_record(a, j + 2, acct, float(j)) for j in range(max(k - len(evs), 0))]
            blobs.append(b"".join(evs)); counts.append(len(evs))
        rec = np.frombuffer(b"".join(blobs), dtype=np.uint8).view(F.REC64)
        off = F.csr_offsets_from_counts(counts)
        model, prog = O.MODEL_BANK_ACCOUNT, P.bank_account_program()
        arrival = S.interleave_arrival(rec, seed=193)
    want, nev, nerr = O.fold_packed(model, O.REC_FIXED64, rec, off)
    with ReplayEngine(0) as e:
        e.register_program(prog)
        e.fold_unsorted(arrival, n_agg)
        assert_same(e.export_states(), want)
        assert (e.stats().n_events, e.stats().n_errors) == (nev, nerr)
        batch = arrival[:777].copy()
        e.fold_incremental(batch)
        assert_same(e.export_states(), O.fold_incremental(model, batch, want))


@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_untouched_instance_with_nan_is_not_changed(kernel):
    """Case-class equals starts with `this eq that`: an account whose balance is NaN and that receives no event in the
    fold keeps its instance and is not CHANGED; one that receives Updated(NaN) is a new instance and is CHANGED
    (tests/test_oracle_golden.py::test_untouched_instance_equals_itself_even_with_nan pins the oracle)."""
    n_agg = 64
    ids = [str(uuid.UUID(int=1000 + i)) for i in range(n_agg)]
    nan = float("nan")
    created = b"".join(F.bank_created_record(i, 1, ids[i], "o", "c", nan if i % 2 else 1.0) for i in range(n_agg))
    off = np.arange(n_agg + 1, dtype=np.uint64) * 64
    ev = np.frombuffer(created, np.uint8)
    init, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, ev, off)
    got, _ = run_engine(P.bank_account_program(), ev, off, kernel=kernel)
    assert_same(got, init, "create")
    # second fold: a third of the accounts get nothing, a third Updated(same value: NaN or 1.0), a third Updated(-0.0)
    recs, counts = [], []
    for i in range(n_agg):
        if i % 3 == 0:
            counts.append(0)
        elif i % 3 == 1:
            recs.append(F.bank_updated_record(i, 2, ids[i], nan if i % 2 else 1.0)); counts.append(1)
        else:
            recs.append(F.bank_updated_record(i, 2, ids[i], -0.0)); counts.append(1)
    ev2 = np.frombuffer(b"".join(recs), np.uint8)
    off2 = np.zeros(n_agg + 1, np.uint64)
    np.cumsum(np.asarray(counts) * 64, out=off2[1:])
    want, _, _ = O.fold_packed(O.MODEL_BANK_ACCOUNT, O.REC_FIXED64, ev2, off2, init)
    flags = want.view(F.BANK_STATE).reshape(-1)["flags"]
    assert all(int(flags[i]) == N.ST_EXISTS for i in range(0, n_agg, 3))                       # untouched, NaN or not
    assert all(int(flags[i]) == (N.ST_EXISTS | (N.ST_CHANGED if i % 2 else 0)) for i in range(1, n_agg, 3))
    got2, _ = run_engine(P.bank_account_program(), ev2, off2, kernel=kernel, init=init)
    assert_same(got2, want, "second fold")


# ------------------------------------------------------------------ raw Kafka record batches -> fold
def test_grow_states_keeps_content():
    n_agg = 3000
    rec, off = S.counter_csr(n_agg, 4, seed=301)
    want, _, _ = O.fold_packed(O.MODEL_COUNTER, O.REC_FIXED64, rec, off)
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        e.load_events(rec, off)
        e.fold()
        e.grow_states(n_agg - 5)       # never shrinks
        assert e.n_aggregates() == n_agg
        e.grow_states(5000)
        got = e.export_states()
        assert got.shape[0] == 5000
        assert_same(got[:n_agg], want, "kept")
        assert not got[n_agg:].any()
        batch = F.counter_records([0, 0, 2], [9, 10, 1], [4999, 4999, 7], [5, 6, 0])
        big = np.zeros((5000, want.shape[1]), np.uint8)
        big[:n_agg] = want
        e.fold_incremental(batch)
        assert_same(e.export_states(), O.fold_incremental(O.MODEL_COUNTER, batch, big), "after growth")
    with ReplayEngine(0) as e:      # from nothing: a table of None
        e.register_program(P.counter_program())
        e.grow_states(10)
        assert not e.export_states().any()


@pytest.mark.parametrize("compression", ["none", "lz4"])
def test_record_batches_to_states(compression):
    """poll -> decode -> fold, several fetches over two partitions with transactions; the state table must equal the
    oracle's fold of the records a read_committed consumer would have delivered (oracle/kafka_batch.py)."""
    import struct

    from oracle import kafka_batch as K
    from surge_b200.ingest import Ingest

    rng = np.random.default_rng(500)
    ing = Ingest()
    nxt = {0: 0, 1: 50}
    want = np.zeros((0, 16), np.uint8)
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        for rnd in range(6):
            fetches = []
            for p in (0, 1):
                buf = bytearray()
                aborted = []
                for _ in range(int(rng.integers(1, 6))):
                    n = int(rng.integers(1, 200))
                    base = nxt[p]
                    recs = []
                    for d in range(n):
                        k = int(rng.integers(0, 400 * (rnd + 1)))
                        t = int(rng.choice([0, 1, 2, 3], p=[0.45, 0.44, 0.1, 0.01]))
                        # the partition is part of the id: all records of an aggregate share a partition, as in Kafka
                        recs.append((d, f"p{p}-agg{k}:{base + d}".encode(), struct.pack("<IIi", t, base + d, int(rng.integers(-2**31, 2**31)))))
                    txn = rng.random() < 0.5
                    pid = int(rng.integers(1, 5))
                    buf += K.encode_record_batch(base, recs, compression=compression, producer_id=pid if txn else -1, transactional=txn)
                    nxt[p] += n
                    if txn:
                        abort = rng.random() < 0.3
                        if abort:
                            aborted.append((pid, base))
                        buf += K.encode_control_batch(nxt[p], pid, K.ABORT if abort else K.COMMIT)
                        nxt[p] += 1
                    if rng.random() < 0.3:
                        buf += K.encode_record_batch(nxt[p], [(0, b"", b"")])   # a producer's flush record
                        nxt[p] += 1
                fetches.append((p, bytes(buf), aborted))
            for p, buf, aborted in fetches:
                ing.set_aborted(p, aborted)
                ing.record_batches(p, buf)
            batch = ing.pending()
            keys = ing.keys()
            e.fold_ingested(ing)
            assert len(ing.pending()) == 0
            for p in (0, 1):
                assert ing.offsets(p) == (nxt[p], nxt[p])
            cap = e.n_aggregates()
            assert cap >= len(keys) and cap >= 1024
            if cap > want.shape[0]:
                grown = np.zeros((cap, 16), np.uint8)
                grown[: want.shape[0]] = want
                want = grown
            want = O.fold_incremental(O.MODEL_COUNTER, batch.view(F.REC64).reshape(-1), want)
            assert_same(e.export_states(), want, f"round {rnd}")
            for i in rng.integers(0, len(keys), 20):
                st = want.view(F.COUNTER_STATE).reshape(-1)[i]
                got = e.get(keys[i])
                if st["flags"] & N.ST_EXISTS:
                    assert np.frombuffer(got, "<i4").tolist() == [int(st["count"]), int(st["version"])]
                else:
                    assert got is None
            assert e.get("never-seen") is None
        # the checker agrees with its own restatement of the whole session only through the decoder tests
        # (tests/test_ingest_cpu.py); here the point is decode -> GPU fold -> get.


def test_state_topic_restore_from_raw_record_batches():
    """Today's rebuild in the reference, from broker bytes: the compacted STATE topic (snapshot per write, null = delete)
    decoded natively and folded with the snapshot-restore program must equal the object-level KTable restatement
    (SurgeStateStoreConsumer.scala:57-76; AggregateStateStoreKafkaStreamsSpec.scala:64-85)."""
    import struct

    from oracle import kafka_batch as K
    from surge_b200.ingest import Ingest

    rng = np.random.default_rng(612)
    ing = Ingest()
    ing.set_null_value_type(1)
    history = []
    off = 0
    with ReplayEngine(0) as e:
        e.register_program(P.counter_snapshot_restore_program())
        for poll in range(4):
            blob = bytearray()
            for _ in range(5):
                recs = []
                for d in range(int(rng.integers(50, 300))):
                    key = f"agg-{int(rng.integers(0, 500))}"
                    if rng.random() < 0.1:
                        recs.append((d, key.encode(), None)); history.append((key, None))
                    else:
                        c, v = int(rng.integers(-2**31, 2**31)), int(rng.integers(0, 1000))
                        recs.append((d, key.encode(), struct.pack("<IIii", 0, 0, c, v))); history.append((key, (c, v)))
                blob += K.encode_record_batch(off, recs, compression="lz4")
                off += len(recs)
            ing.record_batches(0, bytes(blob))
            e.fold_ingested(ing)
            table = M.ktable_restore(history)
            for key in {k for k, _ in history}:
                got = e.get(key)
                want = table.get(key)
                assert (got is None) == (want is None), key
                if want is not None:
                    assert tuple(np.frombuffer(got, "<i4").tolist()) == want
        assert ing.offsets(0) == (off, off)


def test_json_events_of_the_reference_test_model_fold_to_its_golden_states():
    """The reference's own Counter test model writes Json.toJson(evt) to the events topic (core TestBoundedContext.scala:159-161)
    and its specs pin (3,3) -Incr(1, seq 4)-> (4,4) -Incr(1, seq 5)-> (5,5) (PersistentActorSpec.scala:140-145,181) and the
    multilanguage None -> (1,1) -> (2,2) -Decr-> (1,3) (MultilanguageGatewayServiceImplSpec.scala:72-136): JSON bytes in,
    golden states out, nothing re-encoded on the host."""
    import json

    from oracle import kafka_batch as K
    from surge_b200.ingest import Ingest

    cls = "surge.core.TestBoundedContext."
    ing = Ingest()
    ing.set_json_packer("_type", [(cls + "CountIncremented", 0, [("incrementBy", N.JSON_I32, 16), ("sequenceNumber", N.JSON_I32, 4)]),
                                  (cls + "CountDecremented", 1, [("decrementBy", N.JSON_I32, 16), ("sequenceNumber", N.JSON_I32, 4)]),
                                  (cls + "NoOpEvent", 2, [("sequenceNumber", N.JSON_I32, 4)])], unknown_type=3)
    ing.set_value_framing(N.VALUE_JSON)

    def ev(name, agg, seq, **kw):
        return (f"{agg}:{seq}".encode(), json.dumps({"_type": cls + name, "aggregateId": agg, **kw, "sequenceNumber": seq}, separators=(",", ":")).encode())

    events = [ev("CountIncremented", "a", s, incrementBy=1) for s in (1, 2, 3, 4, 5)]
    events += [ev("CountIncremented", "ml", 1, incrementBy=1), ev("CountIncremented", "ml", 2, incrementBy=1), ev("CountDecremented", "ml", 3, decrementBy=1)]
    events += [ev("NoOpEvent", "n", 9), ev("ExceptionThrowingEvent", "x", 1, errorMsg="boom")]
    recs = [(d, k, v) for d, (k, v) in enumerate(events)]
    with ReplayEngine(0) as e:
        e.register_program(P.counter_program())
        ing.record_batches(0, K.encode_record_batch(0, recs[:3], compression="lz4"))
        e.fold_ingested(ing)
        assert np.frombuffer(e.get("a"), "<i4").tolist() == [3, 3]
        ing.record_batches(0, K.encode_record_batch(3, [(d - 3, k, v) for d, k, v in recs[3:]], compression="lz4"))
        e.fold_ingested(ing)
        assert np.frombuffer(e.get("a"), "<i4").tolist() == [5, 5]
        assert np.frombuffer(e.get("ml"), "<i4").tolist() == [1, 3]
        assert np.frombuffer(e.get("n"), "<i4").tolist() == [0, 0]          # NoOp materialises State(id, 0, 0)
        assert e.get("x") is None                                            # the handler threw: no state
        rows = e.export_states().view(F.COUNTER_STATE).reshape(-1)
        assert int(rows[ing.keys().index("x")]["flags"]) & N.ST_ERROR
