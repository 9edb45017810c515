"""A business app behind the reference's multilanguage gRPC protocol, probed over the wire (surge_b200/multilanguage_probe.py +
surge_b200/infer.py): the fold program of the sample app comes out identical to the hand-written one, a non-foldable app is
refused. The fake apps here are in-process grpc servers that parse and answer the raw `BusinessLogicService.HandleEvents` call
(multilanguage-protocol.proto); the sample logic is multilanguage-scala-sdk-sample Main.scala:25-30 with JSON payloads, the way
the sample serialises them."""
import ctypes as C
import json
import struct
from concurrent import futures

import grpc
import pytest

from surge_b200 import infer as INF
from surge_b200 import multilanguage_probe as MP
from surge_b200 import native as N
from surge_b200 import programs as P


def _serve(handle_events):
    """handle_events(state payload or None, [event payloads]) -> state payload or None; raising aborts the call."""
    def behaviour(request: bytes, context):
        aggregate_id, state, events = MP.decode_handle_events_request(request)
        try:
            out = handle_events(state, events)
        except Exception as ex:   # noqa: BLE001
            context.abort(grpc.StatusCode.INTERNAL, str(ex))
        return MP.encode_handle_events_response(aggregate_id, out)

    class Generic(grpc.GenericRpcHandler):
        def service(self, details):
            if details.method == MP.HANDLE_EVENTS:
                return grpc.unary_unary_rpc_method_handler(behaviour, request_deserializer=lambda b: b, response_serializer=lambda b: b)
            return None

    server = grpc.server(futures.ThreadPoolExecutor(max_workers=2))
    server.add_generic_rpc_handlers((Generic(),))
    port = server.add_insecure_port("127.0.0.1:0")
    server.start()
    return server, grpc.insecure_channel(f"127.0.0.1:{port}")


def _jvm_int(x):
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


# adapters of the sample model: state payload {"balance": n}, event payload {"amount": n}
_state_to_payload = lambda packed: json.dumps({"balance": struct.unpack_from("<i", packed, 0)[0]}).encode()                      # noqa: E731
_payload_to_state = lambda payload: struct.pack("<i", json.loads(payload)["balance"]) + bytes(4)                                  # noqa: E731
_record_to_event = lambda rec: json.dumps({"amount": struct.unpack_from("<i", rec, 16)[0]}).encode()                              # noqa: E731


def test_protobuf_framing_round_trips():
    req = MP.encode_handle_events_request("agg-1", b"\x01\x02", [b"e1", b"", b"e3"])
    assert MP.decode_handle_events_request(req) == ("agg-1", b"\x01\x02", [b"e1", b"", b"e3"])
    assert MP.decode_handle_events_request(MP.encode_handle_events_request("a", None, [b"x"])) == ("a", None, [b"x"])
    assert MP.decode_handle_events_response(MP.encode_handle_events_response("a", None)) is None
    assert MP.decode_handle_events_response(MP.encode_handle_events_response("a", b"")) == b""        # Some(empty payload) is not None
    assert MP.decode_handle_events_response(MP.encode_handle_events_response("a", b"s")) == b"s"


def test_sample_business_app_probed_over_grpc_gives_the_hand_written_program():
    def app(state, events):          # Main.scala:25-30, folded over the request's events
        bal = None if state is None else json.loads(state)["balance"]
        for e in events:
            amount = json.loads(e)["amount"]
            bal = amount if bal is None else _jvm_int(bal + amount)
        return None if bal is None else json.dumps({"balance": bal}).encode()

    server, channel = _serve(app)
    try:
        handler = MP.grpc_handler(channel, _state_to_payload, _payload_to_state, _record_to_event)
        got = INF.infer_program(handler, 8, 1, probes=8, check_sequences=30, check_length=10)
        want = P.int_balance_program()
        assert bytes(C.string_at(C.addressof(got.program()), C.sizeof(want))) == bytes(C.string_at(C.addressof(want), C.sizeof(want)))
    finally:
        channel.close()
        server.stop(0)


def test_a_business_app_that_is_not_a_fold_program_is_refused():
    def app(state, events):          # interest: balance * 2 + amount — no word-wise transformer does that
        bal = 0 if state is None else json.loads(state)["balance"]
        for e in events:
            bal = _jvm_int(bal * 2 + json.loads(e)["amount"])
        return json.dumps({"balance": bal}).encode()

    server, channel = _serve(app)
    try:
        handler = MP.grpc_handler(channel, _state_to_payload, _payload_to_state, _record_to_event)
        with pytest.raises(INF.InferenceError):
            INF.infer_program(handler, 8, 1, probes=8, check_sequences=10, check_length=6)
    finally:
        channel.close()
        server.stop(0)


def test_an_app_that_fails_on_one_event_class_gets_a_throw_rule():
    def app(state, events):
        bal = None if state is None else json.loads(state)["balance"]
        for e in events:
            ev = json.loads(e)
            if ev.get("kind") == 1:
                raise ValueError("unsupported event")
            bal = ev["amount"] if bal is None else _jvm_int(bal + ev["amount"])
        return None if bal is None else json.dumps({"balance": bal}).encode()

    rec_to_event = lambda rec: json.dumps({"kind": struct.unpack_from("<I", rec, 0)[0], "amount": struct.unpack_from("<i", rec, 16)[0]}).encode()   # noqa: E731
    server, channel = _serve(app)
    try:
        handler = MP.grpc_handler(channel, _state_to_payload, _payload_to_state, rec_to_event)
        got = INF.infer_program(handler, 8, 2, probes=8, check_sequences=20, check_length=8)
        assert got.rules == [(N.MATERIALISE, [(N.OP_ADD_I32, 0, 16, 4)]), (N.THROW, [])]
    finally:
        channel.close()
        server.stop(0)
