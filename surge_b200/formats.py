"""Packed layouts of the replay path: the binary SurgeAggregateFormatting of the benchmark models.

The reference's state/event byte contract is whatever the model's formatters emit
(modules/serialization/src/main/scala/surge/core/SurgeFormatting.scala:5-17); its samples use
play-json. The batch engine needs fixed layouts, so each sample model gets a little-endian
binary formatting, defined HERE once and restated independently by the oracle:

  record (fixed64)   +0 u32 type  +4 u32 seq  +8 u64 agg  +16 payload[48]
  record (var16)     +0 u32 type  +4 u32 seq  +8 u32 payload_len  +12 u32 agg  +16 payload (padded to 16)
  state              program bytes, then u32 flags, u32 err_idx (engine-owned)

  Counter   (scaladsl TestBoundedContext.scala:13-89)
      event types 0 CountIncremented, 1 CountDecremented, 2 NoOpEvent, 3 ExceptionThrowingEvent
      payload: i32 by @16          state (16 B): i32 count @0, i32 version @4
  BankAccount (surge-docs BankAccountCommandModel.scala:19,39,46)
      event types 0 BankAccountCreated, 1 BankAccountUpdated
      payload: uuid[16] @16, f64 balance @32, owner @40 (u8 len + 15 bytes), code @56 (u8 len + 7 bytes)
      state (64 B): uuid @0, f64 balance @16, owner @24, code @40, zero pad @48
  IntBalance (multilanguage-scala-sdk-sample Main.scala:19-30)
      event type 0 MoneyDeposited; payload i32 amount @16; state (16 B): i32 balance @0
"""
from __future__ import annotations

import json
import struct
import uuid as _uuid
from typing import Optional, Sequence, Tuple

import numpy as np

REC_BYTES = 64

REC64 = np.dtype([("type", "<u4"), ("seq", "<u4"), ("agg", "<u8"), ("arg0", "<i4"), ("arg1", "<i4"),
                  ("arg2", "<f8"), ("pad", "u1", 32)])
assert REC64.itemsize == 64

COUNTER_STATE = np.dtype([("count", "<i4"), ("version", "<i4"), ("flags", "<u4"), ("err_idx", "<u4")])
INT_BALANCE_STATE = np.dtype([("balance", "<i4"), ("pad", "<u4"), ("flags", "<u4"), ("err_idx", "<u4")])
BANK_STATE = np.dtype([("uuid", "u1", 16), ("balance", "<f8"), ("owner", "u1", 16), ("code", "u1", 8),
                       ("pad", "u1", 8), ("flags", "<u4"), ("err_idx", "<u4")])
assert COUNTER_STATE.itemsize == 16 and BANK_STATE.itemsize == 64

# Counter event types
COUNT_INCREMENTED, COUNT_DECREMENTED, NO_OP_EVENT, EXCEPTION_THROWING_EVENT = 0, 1, 2, 3
# BankAccount event types
BANK_ACCOUNT_CREATED, BANK_ACCOUNT_UPDATED = 0, 1
MONEY_DEPOSITED = 0


def counter_records(types, seqs, aggs, bys) -> np.ndarray:
    """Fixed 64-byte Counter records as an (n,) REC64 array."""
    n = len(types)
    r = np.zeros(n, dtype=REC64)
    r["type"] = types
    r["seq"] = seqs
    r["agg"] = aggs
    r["arg0"] = bys
    return r


def csr_offsets_from_counts(counts, rec_bytes: int = REC_BYTES) -> np.ndarray:
    off = np.zeros(len(counts) + 1, dtype=np.uint64)
    np.cumsum(np.asarray(counts, dtype=np.uint64) * np.uint64(rec_bytes), out=off[1:])
    return off


def _pstr(s: str, slot: int) -> bytes:
    b = s.encode("utf-8")
    if len(b) > slot - 1:
        raise ValueError(f"string {s!r} does not fit a {slot}-byte slot")
    return bytes([len(b)]) + b + bytes(slot - 1 - len(b))


def _unpstr(b: bytes) -> str:
    return bytes(b[1:1 + b[0]]).decode("utf-8")


def bank_created_record(agg: int, seq: int, account: str, owner: str, code: str, balance: float) -> bytes:
    return struct.pack("<IIQ", BANK_ACCOUNT_CREATED, seq, agg) + _uuid.UUID(account).bytes + struct.pack("<d", balance) + \
        _pstr(owner, 16) + _pstr(code, 8)


def bank_updated_record(agg: int, seq: int, account: str, new_balance: float) -> bytes:
    return struct.pack("<IIQ", BANK_ACCOUNT_UPDATED, seq, agg) + _uuid.UUID(account).bytes + struct.pack("<d", new_balance) + bytes(24)


def decode_bank_state(row: np.void) -> Optional[dict]:
    if not (int(row["flags"]) & 1):
        return None
    return {"accountNumber": str(_uuid.UUID(bytes=bytes(row["uuid"]))), "accountOwner": _unpstr(bytes(row["owner"])),
            "securityCode": _unpstr(bytes(row["code"])), "balance": float(row["balance"])}


def pack_var_records(types: Sequence[int], seqs: Sequence[int], aggs: Sequence[int], payloads: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    """Variable records -> (bytes u8[total], record byte offsets u64[n+1])."""
    out = bytearray()
    offs = [0]
    for t, s, a, p in zip(types, seqs, aggs, payloads):
        out += struct.pack("<IIII", t, s, len(p), a) + p + bytes((-len(p)) % 16)
        offs.append(len(out))
    return np.frombuffer(bytes(out), dtype=np.uint8).copy(), np.asarray(offs, dtype=np.uint64)


def counter_state_json(aggregate_id: str, count: int, version: int) -> bytes:
    """What the reference's play-json formatting of State(aggregateId,count,version) emits
    (core TestBoundedContext.scala:153): integer-only JSON is predictable — field order is
    case-class order, no whitespace. Used by the host layer to serve JSON models from the
    binary table (the shim calls the user's own writeState on the JVM)."""
    return json.dumps({"aggregateId": aggregate_id, "count": int(count), "version": int(version)},
                      separators=(",", ":"), ensure_ascii=False).encode("utf-8")


# ----------------------------------------------------------------------------- multilanguage protobuf framing
# modules/multilanguage-protocol/src/main/protobuf/multilanguage-protocol.proto:7-20: State, Command and Event are all
#   message X { string aggregateId = 1; bytes payload = 2; }
# and GenericSurgeCommandBusinessLogic.scala:25-38 reads/writes the state and event topics as X.toByteArray.
# proto3 canonical encoding: fields in number order, default (empty) values omitted.
def _pb_uvarint(n: int) -> bytes:
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def multilanguage_proto(aggregate_id: str, payload: bytes) -> bytes:
    """protobuf.State / protobuf.Event .toByteArray for (aggregateId, payload)."""
    aid = aggregate_id.encode("utf-8")
    out = b""
    if aid:
        out += b"\x0a" + _pb_uvarint(len(aid)) + aid
    if payload:
        out += b"\x12" + _pb_uvarint(len(payload)) + bytes(payload)
    return out


def parse_multilanguage_proto(data: bytes):
    """-> (aggregateId, payload); unknown fields are skipped, the last occurrence of a field wins (protobuf semantics)."""
    aid, payload, p = "", b"", 0

    def uvar(p):
        v = shift = 0
        while True:
            b = data[p]
            p += 1
            v |= (b & 0x7F) << shift
            if not b & 0x80:
                return v, p
            shift += 7

    while p < len(data):
        tag, p = uvar(p)
        wt, field = tag & 7, tag >> 3
        if wt == 0:
            _, p = uvar(p)
        elif wt == 1:
            p += 8
        elif wt == 5:
            p += 4
        elif wt == 2:
            ln, p = uvar(p)
            body = data[p:p + ln]
            if len(body) != ln:
                raise ValueError("truncated protobuf field")
            p += ln
            if field == 1:
                aid = body.decode("utf-8")
            elif field == 2:
                payload = bytes(body)
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
    return aid, payload
