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
