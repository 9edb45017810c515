"""Fold programs DERIVED from an event handler by probing it (front-ends that remove hand-written op tables).

A Surge model keeps its `handleEvent` as JVM code — or, in the multilanguage module, as a remote `BusinessLogicService.HandleEvents`
call into a business app written in any language (modules/multilanguage/src/main/scala/com/ukg/surge/multilanguage/
GenericSurgeCommandBusinessLogic.scala:25-38). The GPU cannot call either; it needs the declarative form (include/sgr.h
`sgr_fold_program`). `surge_b200/programs.py` writes those tables by hand, `surge_b200/dsl.py` compiles them from a short text.
This module needs neither: given the handler as a BLACK BOX over the packed forms —

    handler(state: Optional[bytes], record: bytes) -> Optional[bytes]        raises = the handler throws

(the model's own handleEvent between its state codec and its event packer, the same two adapters the store already needs,
shim/scala GpuStateCodec) — it probes the handler with random records and with states the handler itself produced, per event type, and reads off

    the exists-rule   None -> None / Some, Some -> None, result independent of the prior state, always throws
    per state word    kept | copied from a record offset | prior value +/- a record word (32- or 64-bit, wrapping)

then folds random event sequences through both the handler and the derived table and refuses the model unless they agree step
by step. A handler outside the transformer algebra (a product, a branch on the state, a throw that depends on data, a constant
that is not the zero default) is REFUSED with the reason — it keeps the stock RocksDB store, exactly like a model
`sgr_register_program` declines; nothing is ever approximated.

What cannot be seen from bytes is declared by the caller: which state fields are JVM Doubles (they change the publish rule's
comparison, include/sgr.h `f64_field_off`), and the record layout (type @0, seq @4, payload @16..64 for fixed records).
The step-by-step check inside this module is a front-end self-test on a few thousand random events, not a compute path.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import native as N
from .programs import make_program

Handler = Callable[[Optional[bytes], bytes], Optional[bytes]]
Op = Tuple[int, int, int, int]            # (opcode, dst_off, src_off, len)
Rule = Tuple[int, List[Op]]


class InferenceError(ValueError):
    """The handler cannot be expressed as a fold program (or contradicts itself between probes)."""


@dataclass
class Inferred:
    state_bytes: int                      # program bytes + the 8 bytes of flags / err_idx the engine appends
    rules: List[Rule]
    f64_fields: Tuple[int, ...] = ()
    notes: List[str] = field(default_factory=list)

    def program(self, record_kind: int = N.REC_FIXED64) -> N.sgr_fold_program:
        return make_program(self.state_bytes, record_kind, self.rules, self.f64_fields)


_SRC_OFFSETS = (4,) + tuple(range(16, 64, 4))     # seq and the payload; type (+0) selects the rule, agg (+8) belongs to the engine


def _record(rng: np.random.Generator, etype: int) -> bytes:
    r = bytearray(rng.integers(0, 256, size=64, dtype=np.uint8).tobytes())
    struct.pack_into("<I", r, 0, etype)
    struct.pack_into("<Q", r, 8, 0)
    return bytes(r)


def _call(handler: Handler, state: Optional[bytes], rec: bytes):
    """('ok', new state or None) or ('throw', None)."""
    try:
        out = handler(None if state is None else bytes(state), rec)
    except Exception:   # noqa: BLE001 - whatever the handler raises is "the handler throws"
        return "throw", None
    return "ok", (None if out is None else bytes(out))


def _w32(b: bytes, off: int) -> int:
    return struct.unpack_from("<I", b, off)[0]


def _w64(b: bytes, off: int) -> int:
    return struct.unpack_from("<Q", b, off)[0]


def _reachable_states(handler: Handler, n_types: int, user: int, rng: np.random.Generator, want: int) -> List[bytes]:
    """Packed states the handler itself produces from None under random events. Probing with arbitrary bytes would ask the
    handler about states no codec ever writes (padding that is not zero, a string slot with an impossible length)."""
    pool: List[bytes] = []
    seen = set()
    state: Optional[bytes] = None
    for step in range(want * 40):
        kind, out = _call(handler, state, _record(rng, int(rng.integers(0, n_types))))
        if kind == "ok":
            state = out
            if out is not None and len(out) != user:
                raise InferenceError(f"the handler returned {len(out)} state bytes, the model declares {user}")
            if out is not None and out not in seen:
                seen.add(out)
                pool.append(out)
                if len(pool) >= want:
                    break
        if step % 7 == 6:
            state = None      # restart: short histories and long ones
    return pool


def _infer_rule(handler: Handler, etype: int, user: int, rng: np.random.Generator, probes: int, pool: Sequence[bytes]) -> Rule:
    recs = [_record(rng, etype) for _ in range(probes)]
    olds = [pool[int(rng.integers(0, len(pool)))] for _ in range(probes)]
    from_none = [_call(handler, None, r) for r in recs]
    from_some = [_call(handler, s, r) for s, r in zip(olds, recs)]
    kinds = {k for k, _ in from_none + from_some}
    if kinds == {"throw"}:
        return N.THROW, []
    if "throw" in kinds:
        raise InferenceError(f"event type {etype}: the handler throws for some inputs and not for others (a data-dependent throw is not a fold rule)")
    for _, out in from_none + from_some:
        if out is not None and len(out) != user:
            raise InferenceError(f"event type {etype}: the handler returned {len(out)} state bytes, the model declares {user}")
    none_is_none = [out is None for _, out in from_none]
    some_is_none = [out is None for _, out in from_some]
    if all(some_is_none):
        if not all(none_is_none):
            raise InferenceError(f"event type {etype}: deletes an existing state but creates one from None")
        return N.TOMBSTONE, []
    if any(some_is_none) or (any(none_is_none) and not all(none_is_none)):
        raise InferenceError(f"event type {etype}: whether the result exists depends on the data, not only on whether a state existed")
    if all(none_is_none):
        exists_rule, base_olds, outs = N.IF_EXISTS, olds, [out for _, out in from_some]
    else:
        # Some either way: CREATE when the prior state never shows in the result, else MATERIALISE (None = the zero default)
        independent = all(a[1] == b[1] for a, b in zip(from_none, from_some)) and \
            all(_call(handler, olds[(k + 1) % probes], recs[k])[1] == from_some[k][1] for k in range(probes))
        if independent:
            exists_rule, base_olds, outs = N.CREATE, [bytes(user)] * probes, [out for _, out in from_some]
        else:
            exists_rule, base_olds, outs = N.MATERIALISE, olds, [out for _, out in from_some]
    ops: List[Op] = []
    w = 0
    while w < user:
        old = [_w32(s, w) for s in base_olds]
        new = [_w32(o, w) for o in outs]
        if new == old:
            w += 4
            continue
        found: Optional[Op] = None
        # a 64-bit add first: its low half alone looks like a 32-bit add
        if w + 8 <= user and w % 8 == 0:
            old64 = [_w64(s, w) for s in base_olds]
            new64 = [_w64(o, w) for o in outs]
            for src in _SRC_OFFSETS:
                if src + 8 > 64 or src < 16:
                    continue
                rq = [_w64(r, src) for r in recs]
                if exists_rule == N.CREATE:
                    if [(-n) & 0xFFFFFFFFFFFFFFFF for n in new64] == rq:
                        found = (N.OP_SUB_I64, w, src, 8)          # 0 - x on the fresh default (0 + x is a plain copy, found below)
                elif [(n - o) & 0xFFFFFFFFFFFFFFFF for n, o in zip(new64, old64)] == rq:
                    found = (N.OP_ADD_I64, w, src, 8)
                elif [(o - n) & 0xFFFFFFFFFFFFFFFF for n, o in zip(new64, old64)] == rq:
                    found = (N.OP_SUB_I64, w, src, 8)
                if found:
                    break
        for src in (_SRC_OFFSETS if not found else ()):
            rw = [_w32(r, src) for r in recs]
            if new == rw:
                found = (N.OP_SET, w, src, 4)
            elif exists_rule != N.CREATE and [(n - o) & 0xFFFFFFFF for n, o in zip(new, old)] == rw:
                found = (N.OP_ADD_I32, w, src, 4)
            elif exists_rule != N.CREATE and [(o - n) & 0xFFFFFFFF for n, o in zip(new, old)] == rw:
                found = (N.OP_SUB_I32, w, src, 4)
            elif exists_rule == N.CREATE and [(-n) & 0xFFFFFFFF for n in new] == rw:
                found = (N.OP_SUB_I32, w, src, 4)      # 0 - x on the fresh default
            if found:
                break
        if not found:
            raise InferenceError(f"event type {etype}: state bytes {w}..{w + 4} are neither kept, copied from the record, nor the old value "
                                 f"plus / minus a record word (old {old[0]:#x}, new {new[0]:#x})")
        ops.append(found)
        w += found[3]
    # adjacent copies from adjacent record bytes are one op, up to 16 bytes (a UUID, a padded string slot)
    merged: List[Op] = []
    for op in ops:
        if merged and op[0] == N.OP_SET and merged[-1][0] == N.OP_SET and merged[-1][1] + merged[-1][3] == op[1] and merged[-1][2] + merged[-1][3] == op[2] \
                and merged[-1][3] + op[3] <= 16:
            merged[-1] = (N.OP_SET, merged[-1][1], merged[-1][2], merged[-1][3] + op[3])
        else:
            merged.append(op)
    if len(merged) > N.MAX_OPS:
        raise InferenceError(f"event type {etype}: {len(merged)} field ops, the program format holds {N.MAX_OPS}")
    return exists_rule, merged


def apply_rule(rules: Sequence[Rule], user: int, state: Optional[bytes], rec: bytes) -> Tuple[str, Optional[bytes]]:
    """The written semantics of include/sgr.h for ONE event (front-end self-test; the kernels are checked against the oracle's
    interpreter, oracle/program_interp.py, not against this)."""
    etype = _w32(rec, 0)
    if etype >= len(rules) or rules[etype][0] == N.THROW:
        return "throw", None
    exists_rule, ops = rules[etype]
    if exists_rule == N.TOMBSTONE:
        return "ok", None
    if exists_rule == N.IF_EXISTS:
        if state is None:
            return "ok", None
        cur = bytearray(state)
    elif exists_rule == N.MATERIALISE:
        cur = bytearray(state) if state is not None else bytearray(user)
    else:
        cur = bytearray(user)
    for opcode, dst, src, ln in ops:
        if opcode == N.OP_SET:
            cur[dst:dst + ln] = rec[src:src + ln]
        elif opcode in (N.OP_ADD_I32, N.OP_SUB_I32):
            a, b = _w32(cur, dst), _w32(rec, src)
            struct.pack_into("<I", cur, dst, (a + b if opcode == N.OP_ADD_I32 else a - b) & 0xFFFFFFFF)
        else:
            a, b = _w64(cur, dst), _w64(rec, src)
            struct.pack_into("<Q", cur, dst, (a + b if opcode == N.OP_ADD_I64 else a - b) & 0xFFFFFFFFFFFFFFFF)
    return "ok", bytes(cur)


def infer_program(handler: Handler, state_user_bytes: int, n_types: int, *, f64_fields: Sequence[int] = (), seed: int = 0,
                  probes: int = 12, check_sequences: int = 200, check_length: int = 24) -> Inferred:
    """Derive the fold program of `handler` for event types 0 .. n_types-1 (a type >= n_types is a MatchError = throw, as in the
    program format). `state_user_bytes`: the model's packed state without the engine's 8 trailing bytes; a multiple of 4."""
    if state_user_bytes <= 0 or state_user_bytes % 4 or state_user_bytes > 120:
        raise InferenceError("the packed state must be 4 .. 120 bytes, a multiple of 4")
    if not 0 < n_types <= N.MAX_TYPES:
        raise InferenceError(f"1 .. {N.MAX_TYPES} event types")
    rng = np.random.default_rng(seed)
    pool = _reachable_states(handler, n_types, state_user_bytes, rng, max(4 * probes, 32))
    if not pool:
        raise InferenceError("no event type ever creates a state from None: nothing to derive a table from")
    rules = [_infer_rule(handler, t, state_user_bytes, rng, probes, pool) for t in range(n_types)]
    # the whole handler against the whole table, step by step, from None and from states the handler produced
    for q in range(check_sequences):
        state: Optional[bytes] = None if q % 2 == 0 else pool[int(rng.integers(0, len(pool)))]
        for step in range(check_length):
            rec = _record(rng, int(rng.integers(0, n_types)))
            want = _call(handler, state, rec)
            got = apply_rule(rules, state_user_bytes, state, rec)
            if want != got:
                raise InferenceError(f"the derived table and the handler disagree on event type {_w32(rec, 0)} (sequence {q}, step {step}): "
                                     f"the handler is outside the fold-program algebra")
            if want[0] == "ok":
                state = want[1]
    return Inferred(state_user_bytes + 8, rules, tuple(f64_fields))
