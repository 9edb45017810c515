"""The host-side mirror of the reference's plugin / state-store interfaces (surge_b200/store.py).
Reads like the reference's own specs: AggregateStateStoreKafkaStreamsSpec.scala:64-85 (KTable last-write-wins),
PersistentActorSpec's mocked getAggregateBytes (:103-109), the plugin loader (SurgeKafkaStreamsPersistencePlugin.scala:27-50)."""
import numpy as np
import pytest

from surge_b200 import formats as F
from surge_b200 import native as N
from surge_b200 import programs as P
from surge_b200 import store as ST


def test_plugin_loader_reads_the_reference_config_keys():
    ST.GpuReplayPersistencePlugin.program_factory = staticmethod(P.counter_program)
    cfg = {"surge.kafka-streams.state-store-plugin": "gpu-replay", "gpu-replay.plugin-class": "surge_b200.store.GpuReplayPersistencePlugin"}
    plugin = ST.SurgeKafkaStreamsPersistencePluginLoader.load(cfg)
    assert isinstance(plugin, ST.GpuReplayPersistencePlugin) and plugin.enableLogging is False
    sup = plugin.createSupplier("aggregate-state-store")
    assert sup.name() == "aggregate-state-store" and sup.metricsScope() == "gpu-replay"


def test_plugin_loader_fails_loudly_instead_of_falling_back():
    with pytest.raises(KeyError):
        ST.SurgeKafkaStreamsPersistencePluginLoader.load({})
    with pytest.raises(KeyError):
        ST.SurgeKafkaStreamsPersistencePluginLoader.load({"surge.kafka-streams.state-store-plugin": "gpu-replay"})
    with pytest.raises(ModuleNotFoundError):
        ST.SurgeKafkaStreamsPersistencePluginLoader.load({"surge.kafka-streams.state-store-plugin": "x", "x.plugin-class": "no.such.Plugin"})
    with pytest.raises(TypeError):
        ST.SurgeKafkaStreamsPersistencePluginLoader.load({"surge.kafka-streams.state-store-plugin": "x", "x.plugin-class": "collections.OrderedDict"})


def test_record_key_to_aggregate_id():
    assert ST.aggregate_id_of_record_key("agg-1:42") == "agg-1" and ST.aggregate_id_of_record_key("bare") == "bare"


def _ev(t, seq, by):
    return F.counter_records([t], [seq], [0], [by]).tobytes()


@pytest.mark.gpu
def test_store_restores_from_events_and_serves_get_aggregate_bytes():
    ST.GpuReplayPersistencePlugin.program_factory = staticmethod(P.counter_program)
    store = ST.GpuReplayPersistencePlugin().createSupplier("s").get()
    with pytest.raises(ST.InvalidStateStoreException):
        store.get("a")  # not open
    store.init()
    with pytest.raises(N.InvalidStateStoreException):
        store.get("a")  # open but not restored: the reference passes InvalidStateStoreException through as a failed Future
    # multilanguage Counter vector: None -Incr-> (1,1) -Incr-> (2,2) -Decr-> (1,3)   (MultilanguageGatewayServiceImplSpec.scala:72-136)
    store.restore([("a:1", _ev(0, 1, 1)), ("b:1", _ev(0, 1, 5)), ("a:2", _ev(0, 2, 1)), (None, b"\0" * 64), ("a:3", _ev(1, 3, 1))])
    facade = ST.AggregateStateStore(store)
    a = facade.getAggregateBytes("a").result()
    assert np.frombuffer(a, dtype="<i4").tolist() == [1, 3]
    assert np.frombuffer(store.get("b"), dtype="<i4").tolist() == [5, 1]
    assert facade.getAggregateBytes("nope").result() is None
    # later batches append to live aggregates (ApplyEvents on a live actor)
    store.restore([("b:2", _ev(1, 2, 7)), ("c:1", _ev(2, 1, 0))])
    assert np.frombuffer(store.get("b"), dtype="<i4").tolist() == [-2, 2]
    assert np.frombuffer(store.get("c"), dtype="<i4").tolist() == [0, 0]      # NoOp materialises State(id,0,0)
    # many new keys force the table to grow
    store.restore([(f"k{i}:1", _ev(0, 1, i)) for i in range(3000)])
    assert np.frombuffer(store.get("k2999"), dtype="<i4").tolist() == [2999, 1]
    assert np.frombuffer(store.get("a"), dtype="<i4").tolist() == [1, 3]
    assert store.approximateNumEntries() == 3003
    assert facade.healthCheck()["status"] == "up"
    facade.stop()
    store.close()
    assert not store.isOpen()


@pytest.mark.gpu
def test_state_records_keep_ktable_semantics():
    """AggregateStateStoreKafkaStreamsSpec.scala:64-85: put state1(int=1) ... then state1(int=3) => get == the latter;
    a null value deletes (SurgeModel.scala:62-64)."""
    store = ST.GpuReplayKeyValueStore("s", P.counter_program())
    store.init()
    store.restore([])
    js = lambda s, i: ('{"string":"%s","int":%d}' % (s, i)).encode()  # noqa: E731
    store.put("state1", js("state1", 1)); store.put("state2", js("state2", 2)); store.put("state1", js("state1", 3))
    assert store.get("state1") == js("state1", 3) and store.get("state2") == js("state2", 2)
    assert store.putIfAbsent("state2", b"x") == js("state2", 2) and store.get("state2") == js("state2", 2)
    assert store.delete("state1") == js("state1", 3) and store.get("state1") is None
    assert dict(store.all()) == {"state2": js("state2", 2)}
    store.close()


@pytest.mark.gpu
def test_json_model_is_served_through_a_formatter():
    """JSON models: the table is binary, the shim formats on read (play-json bytes of State(aggregateId,count,version))."""
    fmt = lambda key, b: F.counter_state_json(key, *np.frombuffer(b, dtype="<i4").tolist())  # noqa: E731
    store = ST.GpuReplayKeyValueStore("s", P.counter_program(), state_formatter=fmt)
    store.init()
    store.restore([("agg:4", _ev(0, 4, 1)), ("agg:5", _ev(0, 5, 1))])
    assert store.get("agg") == b'{"aggregateId":"agg","count":2,"version":5}'
    store.close()


@pytest.mark.gpu
def test_store_restores_from_raw_record_batches_and_reports_committed_offsets():
    """The whole restore loop on raw broker bytes: fetch -> restore_record_batches -> flush -> getAggregateBytes, and the
    offsets the lag gate compares (KafkaAdminClient.scala:44-56)."""
    import struct

    from oracle import kafka_batch as K

    ev = lambda t, seq, by: struct.pack("<IIi", t, seq, by)  # noqa: E731
    store = ST.GpuReplayKeyValueStore("s", P.counter_program())
    store.init()
    assert store.committed_offsets([0, 1]) == {0: 0, 1: 0}
    p0 = K.encode_record_batch(0, [(0, b"", b"")]) + K.encode_record_batch(1, [(0, b"a:1", ev(0, 1, 1)), (1, b"a:2", ev(0, 2, 1))], compression="lz4",
                                                                          producer_id=5, transactional=True) + K.encode_control_batch(3, 5, K.COMMIT)
    p1 = K.encode_record_batch(10, [(0, b"b:1", ev(0, 1, 9))], producer_id=6, transactional=True) + K.encode_control_batch(11, 6, K.ABORT) + \
        K.encode_record_batch(12, [(0, b"c:1", ev(2, 1, 0))])
    st = store.restore_record_batches(0, p0)
    assert (st["n_records"], st["n_markers"], st["n_control_batches"]) == (2, 1, 1)
    store.restore_record_batches(1, p1, aborted=[(6, 10)])
    assert store.committed_offsets([0, 1]) == {0: 0, 1: 0}          # decoded, not folded yet: the gate must stay shut
    store.flush()
    assert store.committed_offsets([0, 1]) == {0: 4, 1: 13}
    assert np.frombuffer(store.get("a"), dtype="<i4").tolist() == [2, 2]
    assert store.get("b") is None                                    # its only event belonged to an aborted transaction
    assert np.frombuffer(store.get("c"), dtype="<i4").tolist() == [0, 0]
    store.restore_record_batches(0, K.encode_record_batch(4, [(0, b"a:3", ev(1, 3, 1))]))
    store.flush()
    assert np.frombuffer(store.get("a"), dtype="<i4").tolist() == [1, 3]   # the multilanguage Counter vector (1,3)
    assert store.committed_offsets([0]) == {0: 5}
    assert [k for k, _ in store.all()] == ["a", "c"]
    store.put("é", b"1"); store.put("z", b"2"); store.put("B", b"3")
    assert [k for k, _ in store.all()] == ["B", "a", "c", "z", "é"]             # Bytes order: by UTF-8 bytes, not by code point collation
    assert [k for k, _ in store.range("a", "z")] == ["a", "c", "z"]
    with pytest.raises(N.SgrError):
        store.put_event("x:1", bytes(64))
    store.close()


@pytest.mark.gpu
def test_state_topic_records_and_events_fold_into_one_table():
    """Feed (i) + (ii) of the Scala store: put(key, serializedState) / put(key, null) are snapshot / tombstone events of the same
    fold program as the model's own events; flush() folds them in arrival order; get() answers in the model's serialized form.
    Checked against the program interpreter (the written semantics of include/sgr.h)."""
    from oracle import program_interp as I

    codec = ST.StateCodec(
        to_packed=lambda key, b: np.array([int(x) for x in b.decode().split(",")], dtype="<i4").tobytes(),
        from_packed=lambda key, p: ",".join(str(x) for x in np.frombuffer(p, dtype="<i4")).encode(),
        snapshot_type=P.COUNTER_SNAPSHOT_TYPE, tombstone_type=P.COUNTER_TOMBSTONE_TYPE)
    store = ST.GpuReplayKeyValueStore("s", P.counter_program_with_snapshot_rules(), codec=codec)
    store.init()
    store.put("a", b"10,3"); store.put("b", b"7,1"); store.put("", b"")           # the last one is a flush record: ignored
    assert store.get("a") == b"10,3"                                                # read-your-writes before the fold
    store.flush()
    assert store.get("a") == b"10,3" and store.get("b") == b"7,1" and store.get("zz") is None
    store.put_event("a:4", _ev(0, 4, 5))          # an event on top of the snapshot: count 15, version 4
    store.put("b", b"100,9")                      # a newer snapshot of b wins over the old one
    store.put_event("b:10", _ev(1, 10, 1))        # ... and an event after it applies to the NEW snapshot: 99, 10
    store.put("c", b"1,1"); store.delete("c")     # written and deleted inside one batch
    store.flush()
    assert store.get("a") == b"15,4" and store.get("b") == b"99,10" and store.get("c") is None
    assert [k for k, _ in store.all()] == ["a", "b"]
    store.delete("a"); store.flush()
    assert store.get("a") is None and [k for k, _ in store.all()] == ["b"]
    # the same log through the interpreter
    rules = [(I.MATERIALISE, [(I.OP_ADD_I32, 0, 16, 4), (I.OP_SET, 4, 4, 4)]), (I.MATERIALISE, [(I.OP_SUB_I32, 0, 16, 4), (I.OP_SET, 4, 4, 4)]),
             (I.MATERIALISE, []), (I.THROW, []), (I.CREATE, [(I.OP_SET, 0, 16, 4), (I.OP_SET, 4, 20, 4)]), (I.TOMBSTONE, [])]
    snap = lambda c, v: F.counter_records([4], [0], [0], [c]).tobytes()[:20] + np.int32(v).tobytes() + bytes(40)   # noqa: E731
    log_b = np.frombuffer(snap(7, 1) + snap(100, 9) + _ev(1, 10, 1), dtype=np.uint8)
    want_b = I.fold(rules, 16, log_b, [0, len(log_b)])
    assert np.frombuffer(want_b[0, :8].tobytes(), dtype="<i4").tolist() == [99, 10]
    store.close()
