"""The multilanguage business app's `HandleEvents` as the black box of surge_b200/infer.py.

In Surge's multilanguage module the event handler is not JVM code: `GenericAsyncAggregateCommandModel.handleEvents` ships the
state and the events to the business app over gRPC and takes the new state from the reply
(modules/multilanguage/src/main/scala/com/ukg/surge/multilanguage/GenericAsyncAggregateCommandModel.scala:83-103;
service and messages: modules/multilanguage-protocol/src/main/protobuf/multilanguage-protocol.proto —
`BusinessLogicService.HandleEvents(HandleEventsRequest{aggregateId=1, State state=2, repeated Event events=3})
 -> HandleEventsResponse{aggregateId=1, State state=2}`, State / Event = `{string aggregateId=1; bytes payload=2}`, an unset
state = None in both directions). This module speaks exactly that call with hand-encoded protobuf (no generated stubs: the proto
has no package, so the method is `/BusinessLogicService/HandleEvents`) and wraps it as

    handler(packed state or None, 64-byte record) -> packed state or None          (an RPC error = the handler throws)

between the three adapters every GPU-backed model needs anyway: packed state <-> the app's state payload, packed record -> the
app's event payload. `infer.infer_program(handler, ...)` then derives the fold program of a business app written in any language
without reading its source — or refuses it.
"""
from __future__ import annotations

from typing import Callable, Optional

from .formats import multilanguage_proto, parse_multilanguage_proto

HANDLE_EVENTS = "/BusinessLogicService/HandleEvents"


def _uvarint(n: int) -> bytes:
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def _field(number: int, body: bytes) -> bytes:
    return _uvarint((number << 3) | 2) + _uvarint(len(body)) + body


def encode_handle_events_request(aggregate_id: str, state_payload: Optional[bytes], event_payloads) -> bytes:
    out = _field(1, aggregate_id.encode("utf-8")) if aggregate_id else b""
    if state_payload is not None:
        out += _field(2, multilanguage_proto(aggregate_id, state_payload))
    for p in event_payloads:
        out += _field(3, multilanguage_proto(aggregate_id, p))
    return out


def _fields(data: bytes):
    """(field number, wire type, value) of a protobuf message; length-delimited values as bytes."""
    p = 0
    while p < len(data):
        tag = shift = 0
        while True:
            b = data[p]; p += 1
            tag |= (b & 0x7F) << shift
            if not b & 0x80:
                break
            shift += 7
        number, wt = tag >> 3, tag & 7
        if wt == 2:
            ln = shift = 0
            while True:
                b = data[p]; p += 1
                ln |= (b & 0x7F) << shift
                if not b & 0x80:
                    break
                shift += 7
            yield number, wt, data[p:p + ln]
            p += ln
        elif wt == 0:
            while data[p] & 0x80:
                p += 1
            p += 1
            yield number, wt, None
        elif wt == 1:
            p += 8
            yield number, wt, None
        elif wt == 5:
            p += 4
            yield number, wt, None
        else:
            raise ValueError(f"protobuf wire type {wt}")


def decode_handle_events_request(data: bytes):
    """-> (aggregateId, state payload or None, [event payloads]) — the server side of the call (tests, fake business apps)."""
    aggregate_id, state, events = "", None, []
    for number, wt, value in _fields(data):
        if wt != 2:
            continue
        if number == 1:
            aggregate_id = value.decode("utf-8")
        elif number == 2:
            state = parse_multilanguage_proto(value)[1]
        elif number == 3:
            events.append(parse_multilanguage_proto(value)[1])
    return aggregate_id, state, events


def encode_handle_events_response(aggregate_id: str, state_payload: Optional[bytes]) -> bytes:
    out = _field(1, aggregate_id.encode("utf-8")) if aggregate_id else b""
    if state_payload is not None:
        out += _field(2, multilanguage_proto(aggregate_id, state_payload))
    return out


def decode_handle_events_response(data: bytes) -> Optional[bytes]:
    state = None
    for number, wt, value in _fields(data):
        if wt == 2 and number == 2:
            state = parse_multilanguage_proto(value)[1]
    return state


def grpc_handler(channel, state_to_payload: Callable[[bytes], bytes], payload_to_state: Callable[[bytes], bytes],
                 record_to_event_payload: Callable[[bytes], bytes], aggregate_id: str = "probe", timeout: float = 5.0):
    """The business app behind `channel` (a grpc.Channel) as infer.Handler. One event per call: the fold of a list is the fold of
    its elements (CommandModels.scala:25-28), and single steps are what the derivation compares."""
    call = channel.unary_unary(HANDLE_EVENTS, request_serializer=lambda b: b, response_deserializer=lambda b: b)

    def handler(state: Optional[bytes], record: bytes) -> Optional[bytes]:
        req = encode_handle_events_request(aggregate_id, None if state is None else state_to_payload(state), [record_to_event_payload(record)])
        out = decode_handle_events_response(call(req, timeout=timeout))       # grpc.RpcError propagates: "the handler throws"
        return None if out is None else payload_to_state(out)

    return handler
