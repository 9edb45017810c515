"""A small text front-end for fold programs: the model author writes the event handler once more, next to
the Scala one, in a form that compiles to include/sgr.h's sgr_fold_program.

    record fixed64
    state Counter:
        count: i32
        version: i32
    event 0 CountIncremented(by: i32 @16, seq: i32 @4): materialise      # agg.getOrElse(State(id, 0, 0)).copy(...)
        count += by
        version = seq
    event 2 NoOpEvent: materialise                                         # => current
    event 3 ExceptionThrowingEvent: throw

  * `state`: fields in layout order. Types: i32, i64, f64 (a JVM Double: bit-copied, compared with == by the publish rule),
    uuid (16 bytes), bytes[N] (N a multiple of 4). The engine appends its 8 bytes and pads to a multiple of 16.
  * `event <type index> <Name>(<param>: <type> @<byte offset in the record>, ...): <exists rule>` with
    materialise | if_exists | create | tombstone | throw — the five ways a Scala handler treats Option[Agg]
    (surge_b200/programs.py quotes the reference handlers they stand for).
  * statements: `field = param` (same size), `field += param`, `field -= param` (i32 / i64, JVM wrap-around).
    Record offsets count from the start of the packed record: +4 is the sequence number, +16 the first payload byte.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

from . import native as N
from .programs import make_program

_SIZES = {"i32": 4, "i64": 8, "f64": 8, "uuid": 16}
_RULES = {"if_exists": N.IF_EXISTS, "materialise": N.MATERIALISE, "materialize": N.MATERIALISE, "create": N.CREATE,
          "tombstone": N.TOMBSTONE, "throw": N.THROW}


class DslError(ValueError):
    def __init__(self, line_no: int, msg: str):
        super().__init__(f"line {line_no}: {msg}")


def _size_of(ty: str, line_no: int) -> int:
    if ty in _SIZES:
        return _SIZES[ty]
    m = re.fullmatch(r"bytes\[(\d+)\]", ty)
    if m and int(m.group(1)) % 4 == 0 and int(m.group(1)) > 0:
        return int(m.group(1))
    raise DslError(line_no, f"unknown type '{ty}' (i32, i64, f64, uuid, bytes[N] with N a multiple of 4)")


@dataclass
class CompiledModel:
    name: str
    program: N.sgr_fold_program
    state_bytes: int
    record_kind: int
    fields: Dict[str, Tuple[int, str]] = field(default_factory=dict)     # name -> (byte offset in the state, type)
    events: Dict[int, str] = field(default_factory=dict)                 # type index -> event class name


def compile_model(text: str) -> CompiledModel:
    record_kind = N.REC_FIXED64
    state_name = None
    fields: Dict[str, Tuple[int, str]] = {}
    state_off = 0
    rules: Dict[int, Tuple[int, List[Tuple[int, int, int, int]]]] = {}
    events: Dict[int, str] = {}
    section = None          # "state" | ("event", index, params)
    for line_no, raw in enumerate(text.splitlines(), 1):
        line = raw.split("#", 1)[0].rstrip()
        if not line.strip():
            continue
        indented = line[0] in " \t"
        body = line.strip()
        if not indented:
            section = None
            m = re.fullmatch(r"record\s+(fixed64|var16)", body)
            if m:
                record_kind = N.REC_FIXED64 if m.group(1) == "fixed64" else N.REC_VAR16
                continue
            m = re.fullmatch(r"state\s+(\w+)\s*:", body)
            if m:
                if state_name is not None:
                    raise DslError(line_no, "a model has one state")
                state_name, section = m.group(1), "state"
                continue
            m = re.fullmatch(r"event\s+(\d+)\s+(\w+)\s*(?:\((.*)\))?\s*:\s*(\w+)", body)
            if m:
                idx, name, plist, rule = int(m.group(1)), m.group(2), m.group(3) or "", m.group(4).lower()
                if idx >= 16:
                    raise DslError(line_no, "event type index must be below 16")
                if idx in rules:
                    raise DslError(line_no, f"event type {idx} is defined twice")
                if rule not in _RULES:
                    raise DslError(line_no, f"unknown exists rule '{rule}' ({', '.join(sorted(set(_RULES)))})")
                params: Dict[str, Tuple[int, str]] = {}
                for p in [q.strip() for q in plist.split(",") if q.strip()]:
                    pm = re.fullmatch(r"(\w+)\s*:\s*([\w\[\]]+)\s*@\s*(\d+)", p)
                    if not pm:
                        raise DslError(line_no, f"parameter '{p}' is not '<name>: <type> @<record offset>'")
                    off = int(pm.group(3))
                    if off % 4:
                        raise DslError(line_no, f"record offset {off} is not a multiple of 4")
                    _size_of(pm.group(2), line_no)
                    params[pm.group(1)] = (off, pm.group(2))
                rules[idx] = (_RULES[rule], [])
                events[idx] = name
                section = ("event", idx, params)
                continue
            raise DslError(line_no, f"cannot parse '{body}'")
        if section == "state":
            m = re.fullmatch(r"(\w+)\s*:\s*([\w\[\]]+)", body)
            if not m:
                raise DslError(line_no, f"field '{body}' is not '<name>: <type>'")
            ty = m.group(2)
            size = _size_of(ty, line_no)
            if m.group(1) in fields:
                raise DslError(line_no, f"field '{m.group(1)}' is defined twice")
            fields[m.group(1)] = (state_off, ty)
            state_off += size
            continue
        if isinstance(section, tuple):
            _, idx, params = section
            m = re.fullmatch(r"(\w+)\s*(=|\+=|-=)\s*(\w+)", body)
            if not m:
                raise DslError(line_no, f"statement '{body}' is not '<field> (=|+=|-=) <param>'")
            fname, op, pname = m.groups()
            if fname not in fields:
                raise DslError(line_no, f"unknown state field '{fname}'")
            if pname not in params:
                raise DslError(line_no, f"unknown event parameter '{pname}'")
            if rules[idx][0] in (N.TOMBSTONE, N.THROW):
                raise DslError(line_no, "a tombstone / throw rule has no statements")
            (dst, fty), (src, pty) = fields[fname], params[pname]
            fs, ps = _size_of(fty, line_no), _size_of(pty, line_no)
            if op == "=":
                if fs != ps:
                    raise DslError(line_no, f"'{fname}' is {fs} bytes, '{pname}' is {ps}")
                opcode = N.OP_SET
            else:
                if fty != pty or fty not in ("i32", "i64"):
                    raise DslError(line_no, "+= and -= take i32 or i64 on both sides (Doubles are only copied on the replay path)")
                opcode = {("+=", "i32"): N.OP_ADD_I32, ("-=", "i32"): N.OP_SUB_I32, ("+=", "i64"): N.OP_ADD_I64, ("-=", "i64"): N.OP_SUB_I64}[(op, fty)]
            if len(rules[idx][1]) >= 8:
                raise DslError(line_no, "at most 8 statements per event")
            rules[idx][1].append((opcode, dst, src, fs))
            continue
        raise DslError(line_no, f"'{body}' is outside a state or event block")
    if state_name is None:
        raise DslError(0, "no state block")
    if not rules:
        raise DslError(0, "no event blocks")
    n_types = max(rules) + 1
    state_bytes = ((state_off + 8 + 15) // 16) * 16
    if state_bytes > N.MAX_STATE_BYTES:
        raise DslError(0, f"state of {state_off} bytes does not fit {N.MAX_STATE_BYTES - 8}")
    full = [rules.get(t, (N.THROW, [])) for t in range(n_types)]     # an undefined index is a scala.MatchError
    f64 = [off for off, ty in fields.values() if ty == "f64"]
    prog = make_program(state_bytes, record_kind, full, f64_fields=f64)
    return CompiledModel(state_name, prog, state_bytes, record_kind, fields, events)


COUNTER = """
record fixed64
state State:                       # scaladsl TestBoundedContext.scala:28 (aggregateId is the key, not part of the table row)
    count: i32
    version: i32
event 0 CountIncremented(by: i32 @16, seq: i32 @4): materialise
    count += by
    version = seq
event 1 CountDecremented(by: i32 @16, seq: i32 @4): materialise
    count -= by
    version = seq
event 2 NoOpEvent: materialise
event 3 ExceptionThrowingEvent: throw
"""

BANK_ACCOUNT = """
record fixed64
state BankAccount:                 # surge-docs BankAccountCommandModel.scala:19
    accountNumber: uuid
    balance: f64
    accountOwner: bytes[16]
    securityCode: bytes[8]
event 0 BankAccountCreated(accountNumber: uuid @16, balance: f64 @32, accountOwner: bytes[16] @40, securityCode: bytes[8] @56): create
    accountNumber = accountNumber
    balance = balance
    accountOwner = accountOwner
    securityCode = securityCode
event 1 BankAccountUpdated(newBalance: f64 @32): if_exists
    balance = newBalance
"""

INT_BALANCE = """
record fixed64
state BankAccount:                 # multilanguage-scala-sdk-sample Main.scala:19
    balance: i32
event 0 MoneyDeposited(amount: i32 @16): materialise
    balance += amount
"""
