"""Fold programs derived by probing a handler (surge_b200/infer.py) against the reference's sample handlers.

The black boxes are the object-level restatements of the Scala handlers (oracle/surge_model.py, each citing its source) between
a packer and a codec, i.e. what a maintainer would wrap around the JVM handleEvent. Claims checked here, all on the CPU:
  * Counter, multilanguage Counter, IntBalance and the state-topic snapshot rules come out BYTE-IDENTICAL to the hand-written
    tables of surge_b200/programs.py (so everything the GPU tests prove for those tables holds for the derived ones);
  * BankAccount comes out equivalent (the copies are cut differently): equal under the oracle's program interpreter on random
    sequences that include NaN and -0.0 balances;
  * handlers outside the algebra are refused with a reason, never approximated.
"""
import ctypes as C
import struct
import uuid

import numpy as np
import pytest

from oracle import program_interp as I
from oracle import surge_model as M
from surge_b200 import formats as F
from surge_b200 import infer as INF
from surge_b200 import native as N
from surge_b200 import programs as P


def _bytes_of(prog) -> bytes:
    return bytes(C.string_at(C.addressof(prog), C.sizeof(prog)))


def _canon(rules):
    """Copies cut into 4-byte words: two tables with the same canonical form are the same function (a copy of L bytes IS L/4 word
    copies; adds are already one word or one pair)."""
    out = []
    for exists_rule, ops in rules:
        words = []
        for opcode, dst, src, ln in ops:
            if opcode == N.OP_SET:
                words += [(N.OP_SET, dst + k, src + k, 4) for k in range(0, ln, 4)]
            else:
                words.append((opcode, dst, src, ln))
        out.append((exists_rule, sorted(words, key=lambda o: o[1])))
    return out


# ---------------------------------------------------------------- black boxes: packer -> Scala-restated handler -> codec
def counter_box(handle):
    def h(state, rec):
        etype, seq = struct.unpack_from("<II", rec, 0)
        (by,) = struct.unpack_from("<i", rec, 16)
        seq_i = M.jvm_int(seq)
        evt = {0: lambda: M.CountIncremented("a", by, seq_i), 1: lambda: M.CountDecremented("a", by, seq_i), 2: lambda: M.NoOpEvent("a", seq_i),
               3: lambda: M.ExceptionThrowingEvent("a", seq_i, RuntimeError("boom"))}.get(etype, lambda: object())()
        agg = None if state is None else M.State("a", *struct.unpack("<ii", state))
        out = handle(agg, evt)
        return None if out is None else struct.pack("<ii", out.count, out.version)
    return h


def int_balance_box(state, rec):
    (amount,) = struct.unpack_from("<i", rec, 16)
    agg = None if state is None else M.IntBankAccount(struct.unpack_from("<i", state, 0)[0])
    out = M.int_balance_event_handler(agg, M.MoneyDeposited(amount))
    return struct.pack("<i", out.balance) + bytes(4)


def bank_box(state, rec):
    """The handler only MOVES fields (BankAccountCommandModel.scala:81-86), so the box keeps them as the bytes they are: a box that
    decoded strings would have to reject random bytes, which is the packer's job, not the fold's."""
    etype = struct.unpack_from("<I", rec, 0)[0]
    as_obj = lambda b: M.BankAccount(b[0:16], b[24:40], b[40:48], b[16:24])   # noqa: E731  (uuid, owner, code, balance bits)
    agg = None if state is None else as_obj(state)
    if etype == 0:
        evt = M.BankAccountCreated(rec[16:32], rec[40:56], rec[56:64], rec[32:40])
    elif etype == 1:
        evt = M.BankAccountUpdated(rec[16:32], rec[32:40])
    else:
        evt = object()
    out = M.bank_account_handle_event(agg, evt)
    return None if out is None else out.accountNumber + out.balance + out.accountOwner + out.securityCode + bytes(8)


def snapshot_box(state, rec):
    """KTable restore as a handler: a snapshot record replaces the state, a null value deletes it (SurgeStateStoreConsumer.scala:57-76)."""
    etype = struct.unpack_from("<I", rec, 0)[0]
    if etype == 0:
        return rec[16:24]
    if etype == 1:
        return None
    raise ValueError("scala.MatchError")


# ---------------------------------------------------------------- derived == hand-written
def test_counter_tables_are_derived_byte_for_byte():
    got = INF.infer_program(counter_box(M.counter_handle_event), 8, 4)
    assert got.rules == [(N.MATERIALISE, [(N.OP_ADD_I32, 0, 16, 4), (N.OP_SET, 4, 4, 4)]), (N.MATERIALISE, [(N.OP_SUB_I32, 0, 16, 4), (N.OP_SET, 4, 4, 4)]),
                         (N.MATERIALISE, []), (N.THROW, [])]
    assert _bytes_of(got.program()) == _bytes_of(P.counter_program())
    ml = INF.infer_program(counter_box(M.ml_counter_apply_event), 8, 2)
    assert _bytes_of(ml.program()) == _bytes_of(P.ml_counter_program())


def test_int_balance_and_snapshot_rules_are_derived_byte_for_byte():
    assert _bytes_of(INF.infer_program(int_balance_box, 8, 1).program()) == _bytes_of(P.int_balance_program())
    snap = INF.infer_program(snapshot_box, 8, 2)
    assert _canon(snap.rules) == _canon([(N.CREATE, [(N.OP_SET, 0, 16, 4), (N.OP_SET, 4, 20, 4)]), (N.TOMBSTONE, [])])    # programs.counter_snapshot_restore_program


def test_bank_account_is_derived_equivalent_including_nan_and_negative_zero():
    got = INF.infer_program(bank_box, 56, 2, f64_fields=[16])
    assert got.rules[0][0] == N.CREATE and got.rules[1] == (N.IF_EXISTS, [(N.OP_SET, 16, 32, 8)])
    hand = [(N.CREATE, [(N.OP_SET, 0, 16, 16), (N.OP_SET, 16, 32, 8), (N.OP_SET, 24, 40, 16), (N.OP_SET, 40, 56, 8)]), (N.IF_EXISTS, [(N.OP_SET, 16, 32, 8)])]
    assert _canon(got.rules) == _canon(hand)                     # programs.bank_account_program, cut differently
    assert all(ln <= 16 for _, ops in got.rules for _, _, _, ln in ops)
    rng = np.random.default_rng(3)
    specials = [float("nan"), -0.0, 0.0, 1000.25, float("inf")]
    for trial in range(60):
        n = int(rng.integers(1, 12))
        recs = []
        for k in range(n):
            bal = specials[int(rng.integers(0, len(specials)))]
            if rng.random() < 0.4:
                recs.append(F.bank_created_record(0, k, str(uuid.UUID(int=int(rng.integers(1, 1 << 62)))), "owner", "code", bal))
            else:
                recs.append(F.bank_updated_record(0, k, str(uuid.UUID(int=7)), bal))
        log = np.frombuffer(b"".join(recs), dtype=np.uint8).reshape(-1, 64)
        a = I.fold(hand, 64, log, [0, 64 * n], f64_fields=[16])
        b = I.fold(got.rules, 64, log, [0, 64 * n], f64_fields=[16])
        assert np.array_equal(a, b), trial


# ---------------------------------------------------------------- refusals
@pytest.mark.parametrize("name,box,why", [
    ("product", lambda s, r: struct.pack("<I", (struct.unpack_from("<I", s or bytes(4), 0)[0] * 3 + struct.unpack_from("<I", r, 16)[0]) & 0xFFFFFFFF) + bytes(4), "neither kept"),
    ("constant", lambda s, r: struct.pack("<II", 42, 0), "neither kept"),
    ("data-dependent throw", lambda s, r: (_ for _ in ()).throw(ValueError()) if r[16] & 1 else bytes(8), "throws for some inputs"),
    ("data-dependent delete", lambda s, r: None if r[16] & 1 else bytes(8), "depends on the data"),
    ("branch on the state", lambda s, r: (s if s and s[0] & 1 else r[16:24]), None),
    ("wrong size", lambda s, r: bytes(12), "state bytes"),
])
def test_handlers_outside_the_algebra_are_refused(name, box, why):
    with pytest.raises(INF.InferenceError) as ei:
        INF.infer_program(box, 8, 1)
    if why:
        assert why in str(ei.value), (name, str(ei.value))


def test_sixty_four_bit_adds_and_byte_copies():
    def box(state, rec):
        cur = state if state is not None else bytes(24)
        total = (struct.unpack_from("<q", cur, 0)[0] + struct.unpack_from("<q", rec, 24)[0]) & 0xFFFFFFFFFFFFFFFF
        spent = (struct.unpack_from("<Q", cur, 8)[0] - struct.unpack_from("<Q", rec, 32)[0]) & 0xFFFFFFFFFFFFFFFF
        return struct.pack("<QQ", total, spent) + rec[40:48]
    got = INF.infer_program(box, 24, 1)
    assert got.rules == [(N.MATERIALISE, [(N.OP_ADD_I64, 0, 24, 8), (N.OP_SUB_I64, 8, 32, 8), (N.OP_SET, 16, 40, 8)])]


# ---------------------------------------------------------------- random programs: derive them back from their own interpreter
def _random_rules(rng, user, n_types):
    rules = []
    for _ in range(n_types):
        exists_rule = int(rng.choice([N.IF_EXISTS, N.MATERIALISE, N.MATERIALISE, N.CREATE, N.CREATE, N.TOMBSTONE, N.THROW]))
        ops = []
        if exists_rule in (N.IF_EXISTS, N.MATERIALISE, N.CREATE):
            w = 0
            while w < user and len(ops) < N.MAX_OPS:
                kind = rng.random()
                if kind < 0.35:
                    w += 4                                          # this word is kept
                    continue
                src32 = int(rng.choice([4] + list(range(16, 64, 4))))
                if kind < 0.6:
                    ln = int(rng.choice([4, 4, 8, 16]))
                    ln = 4 if src32 == 4 else min(ln, user - w, 64 - src32)    # (+8..16 is the engine's aggregate index: no handler sees it)
                    ops.append((N.OP_SET, w, src32, ln))
                    w += ln
                elif kind < 0.85 or w % 8 or w + 8 > user:
                    ops.append((int(rng.choice([N.OP_ADD_I32, N.OP_SUB_I32])), w, src32, 4))
                    w += 4
                else:
                    src64 = int(rng.choice(list(range(16, 60, 4))))
                    ops.append((int(rng.choice([N.OP_ADD_I64, N.OP_SUB_I64])), w, src64, 8))
                    w += 8
        rules.append((exists_rule, ops))
    return rules


def _box_of(rules, user):
    def h(state, rec):
        out = I._handle(rules, user, None if state is None else bytearray(state), rec)     # raises on THROW / MatchError
        return None if out is None else bytes(out)
    return h


def test_random_programs_are_derived_back_from_their_interpreter():
    rng = np.random.default_rng(2026)
    derived = refused = 0
    for trial in range(120):
        user = int(rng.choice([8, 8, 16, 24, 56]))
        n_types = int(rng.integers(1, 6))
        rules = _random_rules(rng, user, n_types)
        box = _box_of(rules, user)
        creates = any(r in (N.MATERIALISE, N.CREATE) for r, _ in rules)
        try:
            got = INF.infer_program(box, user, n_types, seed=trial, probes=10, check_sequences=40, check_length=12)
        except INF.InferenceError as ex:
            assert not creates, (trial, rules, str(ex))       # the only legitimate refusal here: nothing ever creates a state
            refused += 1
            continue
        derived += 1
        # same function on every history from None: compare the two tables under the oracle's interpreter
        for q in range(30):
            a = b = None
            for step in range(14):
                rec = bytearray(rng.integers(0, 256, size=64, dtype=np.uint8).tobytes())
                struct.pack_into("<I", rec, 0, int(rng.integers(0, n_types + 1)))             # incl. a MatchError type now and then
                rec = bytes(rec)
                try:
                    na = I._handle(rules, user, a, rec)
                except I._Throw:
                    na = "throw"
                try:
                    nb = I._handle(got.rules, user, b, rec)
                except I._Throw:
                    nb = "throw"
                assert (na == "throw") == (nb == "throw"), (trial, rules, got.rules)
                if na != "throw":
                    assert (na is None) == (nb is None) and (na is None or bytes(na) == bytes(nb)), (trial, step, rules, got.rules)
                    a, b = na, nb
    assert derived >= 80 and refused <= 40, (derived, refused)
