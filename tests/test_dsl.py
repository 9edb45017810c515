"""The text front-end for fold programs (surge_b200/dsl.py) compiles the sample models to exactly the op tables
written by hand in surge_b200/programs.py — which the parity tests pin to the oracle's restatement of the Scala handlers."""
import ctypes as C

import pytest

from surge_b200 import dsl as D
from surge_b200 import native as N
from surge_b200 import programs as P


def raw(prog) -> bytes:
    return bytes((C.c_char * C.sizeof(prog)).from_buffer_copy(prog))


@pytest.mark.parametrize("text,hand", [(D.COUNTER, P.counter_program), (D.BANK_ACCOUNT, P.bank_account_program), (D.INT_BALANCE, P.int_balance_program)])
def test_sample_models_compile_to_the_hand_written_tables(text, hand):
    m = D.compile_model(text)
    assert raw(m.program) == raw(hand())


def test_layout_and_metadata():
    m = D.compile_model(D.BANK_ACCOUNT)
    assert m.name == "BankAccount" and m.state_bytes == 64 and m.record_kind == N.REC_FIXED64
    assert m.fields == {"accountNumber": (0, "uuid"), "balance": (16, "f64"), "accountOwner": (24, "bytes[16]"), "securityCode": (40, "bytes[8]")}
    assert m.events == {0: "BankAccountCreated", 1: "BankAccountUpdated"}
    assert m.program.n_f64_fields == 1 and m.program.f64_field_off[0] == 16
    v = D.compile_model(D.COUNTER.replace("record fixed64", "record var16"))
    assert raw(v.program) == raw(P.counter_program(N.REC_VAR16))


def test_gaps_in_event_indices_are_match_errors_and_i64_ops():
    m = D.compile_model("""
state S:
    total: i64
    last: i32
event 0 Add(v: i64 @16, s: i32 @4): materialise
    total += v
    last = s
event 2 Reset: tombstone
""")
    p = m.program
    assert p.n_types == 3 and p.rules[1].exists_rule == N.THROW and p.rules[2].exists_rule == N.TOMBSTONE
    assert (p.rules[0].ops[0].opcode, p.rules[0].ops[0].dst_off, p.rules[0].ops[0].src_off, p.rules[0].ops[0].len) == (N.OP_ADD_I64, 0, 16, 8)
    assert m.state_bytes == 32


@pytest.mark.parametrize("text,fragment", [
    ("state S:\n    a: i32\nevent 0 E(x: f64 @16): create\n    a = x\n", "'a' is 4 bytes, 'x' is 8"),
    ("state S:\n    a: f64\nevent 0 E(x: f64 @16): create\n    a += x\n", "+= and -= take i32 or i64"),
    ("state S:\n    a: i32\nevent 0 E(x: i32 @18): create\n    a = x\n", "not a multiple of 4"),
    ("state S:\n    a: i32\nevent 0 E(x: i32 @16): explode\n", "unknown exists rule"),
    ("state S:\n    a: i32\nevent 0 E(x: i32 @16): throw\n    a = x\n", "has no statements"),
    ("state S:\n    a: i32\nevent 0 E: create\n    b = x\n", "unknown state field"),
    ("state S:\n    a: i32\nevent 0 E: create\nevent 0 F: create\n", "defined twice"),
    ("state S:\n    a: bytes[130]\nevent 0 E: create\n", "unknown type"),
    ("event 0 E: create\n", "no state block"),
    ("state S:\n    a: bytes[124]\nevent 0 E: create\n", "does not fit"),
])
def test_errors_name_the_line_and_the_reason(text, fragment):
    with pytest.raises(D.DslError) as ei:
        D.compile_model(text)
    assert fragment in str(ei.value)
