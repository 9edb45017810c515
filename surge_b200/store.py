"""Host-side mirror of the reference's state-store boundary, over the C ABI.

Reference interfaces mirrored here (same names, argument meaning and error behaviour; paths relative to the
reference checkout, COMMON = modules/common/src/main/scala/surge):

  SurgeKafkaStreamsPersistencePlugin { createSupplier(storeName); enableLogging }
      COMMON/kafka/streams/SurgeKafkaStreamsPersistencePlugin.scala:12-15
  SurgeKafkaStreamsPersistencePluginLoader.load(config)                       same file, :27-50
  KeyValueBytesStoreSupplier.get(): KeyValueStore[Bytes, Array[Byte]]         (Kafka Streams; in-tree example
      modules/common/src/test/scala/surge/kafka/streams/SingleExceptionThrowingKeyValueStore.scala:18-91)
  AggregateStateStoreKafkaStreams.getAggregateBytes(aggregateId): Future[Option[Array[Byte]]]
      COMMON/kafka/streams/AggregateStateStoreKafkaStreams.scala:83-85
  ThreadPools.ioBoundContext (32 threads)                                     COMMON/kafka/streams/ThreadPools.scala:9-11

The JVM shim (shim/scala) is the real drop-in; this module is its executable twin for the toolchain this
image has, and is what the parity tests drive. It adds nothing to the data path: every fold runs in
libsgr.so on the GPU.
"""
from __future__ import annotations

import importlib
import threading
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import native as N
from .engine import ReplayEngine
from .ingest import Ingest

STATE_STORE_PLUGIN_KEY = "surge.kafka-streams.state-store-plugin"


class InvalidStateStoreException(N.InvalidStateStoreException):
    pass


# --------------------------------------------------------------------------- plugin + loader
class SurgeKafkaStreamsPersistencePlugin:
    """trait SurgeKafkaStreamsPersistencePlugin (SurgeKafkaStreamsPersistencePlugin.scala:12-15)."""

    def createSupplier(self, storeName: str) -> "KeyValueBytesStoreSupplier":  # noqa: N802,N803 - reference names
        raise NotImplementedError

    @property
    def enableLogging(self) -> bool:  # noqa: N802
        raise NotImplementedError


class SurgeKafkaStreamsPersistencePluginLoader:
    """object SurgeKafkaStreamsPersistencePluginLoader (same file, :27-50): reads
    `surge.kafka-streams.state-store-plugin`, then `<name>.plugin-class`, instantiates it with the no-arg
    constructor. DEVIATION (deliberate): the reference silently falls back to RocksDB on any failure
    (:34-39,45-47); a GPU store that silently turns into RocksDB would void every measurement, so this loader
    raises instead."""

    @staticmethod
    def load(config: Dict[str, str]) -> SurgeKafkaStreamsPersistencePlugin:
        name = config.get(STATE_STORE_PLUGIN_KEY)
        if not name:
            raise KeyError(f"{STATE_STORE_PLUGIN_KEY} is not set")
        cls_path = config.get(f"{name}.plugin-class")
        if not cls_path:
            raise KeyError(f"{name}.plugin-class is not set")
        mod, _, cls = cls_path.rpartition(".")
        plugin = getattr(importlib.import_module(mod), cls)()
        if not isinstance(plugin, SurgeKafkaStreamsPersistencePlugin):
            raise TypeError(f"{cls_path} is not a SurgeKafkaStreamsPersistencePlugin")
        return plugin


class KeyValueBytesStoreSupplier:
    def __init__(self, name: str, factory: Callable[[str], "GpuReplayKeyValueStore"]):
        self._name, self._factory = name, factory

    def name(self) -> str:
        return self._name

    def get(self) -> "GpuReplayKeyValueStore":
        return self._factory(self._name)

    def metricsScope(self) -> str:  # noqa: N802
        return "gpu-replay"


class GpuReplayPersistencePlugin(SurgeKafkaStreamsPersistencePlugin):
    """Config-selected drop-in:  surge.kafka-streams.state-store-plugin = "gpu-replay"
                                 gpu-replay.plugin-class = "surge_b200.store.GpuReplayPersistencePlugin"
    enableLogging = false: Kafka Streams must not restore this store from a changelog — it rebuilds from the
    events topic (SurgeStateStoreConsumer.scala:63-75)."""

    program_factory: Optional[Callable[[], N.sgr_fold_program]] = None  # set by the model registration
    device: int = 0

    def createSupplier(self, storeName: str) -> KeyValueBytesStoreSupplier:  # noqa: N802,N803
        if GpuReplayPersistencePlugin.program_factory is None:
            raise N.SgrError(N.SGR_ERR_NO_PROGRAM, "no fold program registered for the model (GpuReplayPersistencePlugin.program_factory)")
        return KeyValueBytesStoreSupplier(storeName, lambda n: GpuReplayKeyValueStore(n, GpuReplayPersistencePlugin.program_factory(),
                                                                                      GpuReplayPersistencePlugin.device))

    @property
    def enableLogging(self) -> bool:  # noqa: N802
        return False


# --------------------------------------------------------------------------- the store
def aggregate_id_of_record_key(key: str) -> str:
    """Event record keys are model-defined ("<aggId>:<seq>" in core TestBoundedContext.scala:160, or the bare id);
    the aggregate is key.takeWhile(_ != ':') under the default partitioner (KafkaPartitioner.scala:38-42)."""
    i = key.find(":")
    return key if i < 0 else key[:i]


class StateCodec:
    """serialized state bytes (what the state topic and the actors hold) <-> the packed program bytes of the GPU table
    (trait GpuStateCodec of shim/scala GpuReplayPersistencePlugin.scala). snapshot_type / tombstone_type are the two extra rules of
    the registered fold program (programs.counter_program_with_snapshot_rules)."""

    def __init__(self, to_packed: Callable[[str, bytes], bytes], from_packed: Callable[[str, bytes], bytes], snapshot_type: int, tombstone_type: int):
        self.to_packed, self.from_packed, self.snapshot_type, self.tombstone_type = to_packed, from_packed, snapshot_type, tombstone_type


class GpuReplayKeyValueStore:
    """KeyValueStore[Bytes, Array[Byte]] whose content is the GPU-folded state table.

    write side (one stream thread in the reference): restore()/put_event() batch packed 64-byte event records;
    flush() folds the batch on the GPU (first batch: group + full fold, later batches: incremental fold).
    put()/delete() of *state* records keep KTable semantics — last write wins per key, null deletes
    (SurgeStateStoreConsumer.scala:57-76) — as an overlay over the folded table.
    read side (32-thread pool in the reference): get() is thread-safe.
    """

    def __init__(self, name: str, program: N.sgr_fold_program, device: int = 0, state_formatter: Optional[Callable[[str, bytes], bytes]] = None,
                 codec: Optional[StateCodec] = None):
        self._name = name
        # with a codec, put()/delete() are records of the STATE topic folded on the GPU as snapshot / tombstone events (feed (i) of
        # the Scala store): flush() — which Kafka Streams calls before it commits offsets — makes them readable from the table
        self._codec = codec
        self._unflushed: Dict[str, Optional[bytes]] = {}
        self._engine = ReplayEngine(device)
        self._engine.register_program(program)
        self._formatter = state_formatter
        self._keys: List[str] = []
        self._index: Dict[str, int] = {}
        self._pending: List[np.ndarray] = []
        self._overlay: Dict[str, Optional[bytes]] = {}
        self._folded = False
        self._capacity = 0
        self._open = False
        self._lock = threading.RLock()
        self._restore_callback = None
        self._keys_loaded = (-1, -1)
        self._ingest: Optional[Ingest] = None  # set once the store is fed raw record batches

    # -- lifecycle (StateStore)
    def name(self) -> str:
        return self._name

    def init(self, context=None, root=None) -> None:
        """Registers the restore callback the way the in-tree example does (context.register(root, (k, v) => ...),
        SingleExceptionThrowingKeyValueStore.scala:84-86)."""
        self._open = True
        self._restore_callback = lambda key, value: self.put_event(key, value)
        if context is not None and hasattr(context, "register"):
            context.register(root, self._restore_callback)

    def persistent(self) -> bool:
        return False

    def isOpen(self) -> bool:  # noqa: N802
        return self._open

    def close(self) -> None:
        with self._lock:
            self._open = False
            self._engine.close()

    # -- event ingestion
    def _slot(self, aggregate_id: str) -> int:
        i = self._index.get(aggregate_id)
        if i is None:
            i = len(self._keys)
            self._index[aggregate_id] = i
            self._keys.append(aggregate_id)
        return i

    def put_event(self, record_key: Optional[str], packed_event: bytes) -> None:
        """One record of the events topic: key -> aggregate id, value = the model's packed 64-byte event."""
        if not record_key:  # the producer's empty-key flush markers (KafkaProducerActorImpl.scala:321-329) are dropped
            return
        if len(packed_event) != 64:
            raise ValueError("packed events are 64 bytes")
        with self._lock:
            if self._ingest is not None:
                raise N.SgrError(N.SGR_ERR_INVALID, "this store is already fed through restore_record_batches")
            rec = np.frombuffer(packed_event, dtype=np.uint8).copy()
            agg_id = aggregate_id_of_record_key(record_key)
            rec[8:16] = np.frombuffer(np.uint64(self._slot(agg_id)).tobytes(), dtype=np.uint8)
            self._pending.append(rec)
            self._unflushed.pop(agg_id, None)   # an event supersedes an unflushed snapshot view

    def restore(self, records: Iterable[Tuple[Optional[str], bytes]]) -> None:
        for k, v in records:
            self.put_event(k, v)
        self.flush()

    def restore_record_batches(self, partition: int, data: bytes, aborted: Sequence[Tuple[int, int]] = ()) -> Dict[str, int]:
        """Raw bytes of one fetch response for `partition` (a concatenation of Kafka RecordBatch v2), plus the
        response's aborted transactions [(producerId, firstOffset)]: decoded natively as a read_committed consumer
        would (SurgeStateStoreConsumer.scala:38) into the pending batch. flush() folds it. A store is fed either this
        way or through put_event, not both (each keeps its own id dictionary)."""
        with self._lock:
            if self._keys:
                raise N.SgrError(N.SGR_ERR_INVALID, "this store is already fed through put_event")
            if self._ingest is None:
                self._ingest = Ingest()
            self._ingest.set_aborted(partition, aborted)
            return self._ingest.record_batches(partition, data)

    def committed_offsets(self, partitions: Iterable[int]) -> Dict[int, int]:
        """Per partition, the offset below which every record is inside the state table: what the consumer acting for
        this store commits for the streams application id, so that the producer's lag check
        (KafkaProducerActorImpl.scala:684-708 -> KafkaAdminClient.consumerLag, KafkaAdminClient.scala:44-56) reaches zero
        exactly when get() can serve the state."""
        with self._lock:
            if self._ingest is None:
                return {int(p): 0 for p in partitions}
            return {int(p): self._ingest.offsets(int(p))[1] for p in partitions}

    def flush(self) -> None:
        with self._lock:
            if self._ingest is not None:
                self._engine.fold_ingested(self._ingest)
                self._folded = True
                return
            if not self._pending and self._folded:
                return
            batch = np.concatenate(self._pending) if self._pending else np.zeros(0, dtype=np.uint8)
            self._pending = []
            n_keys = len(self._keys)
            if not self._folded or n_keys > self._capacity:
                # (re)build: carry the current table into a larger one, then append the batch
                prior = None
                if self._folded:
                    old = self._engine.export_states()
                    self._capacity = max(2 * n_keys, 1024)
                    prior = np.zeros((self._capacity, self._engine.state_bytes), dtype=np.uint8)
                    prior[: len(old)] = old
                else:
                    self._capacity = max(2 * n_keys, 1024)
                    prior = np.zeros((self._capacity, self._engine.state_bytes), dtype=np.uint8)
                self._engine.set_initial_states(prior)
                self._folded = True
            self._engine.fold_incremental(batch)
            if self._keys_loaded != (n_keys, self._capacity):
                self._engine.load_keys(self._keys + [f"\0unused-{i}" for i in range(n_keys, self._capacity)])
                self._keys_loaded = (n_keys, self._capacity)
            self._unflushed.clear()

    # -- KeyValueStore
    def _state_record(self, key: str, value: Optional[bytes]) -> None:
        rec = np.zeros(64, dtype=np.uint8)
        rec[0:4] = np.frombuffer(np.uint32(self._codec.tombstone_type if value is None else self._codec.snapshot_type).tobytes(), dtype=np.uint8)
        rec[8:16] = np.frombuffer(np.uint64(self._slot(key)).tobytes(), dtype=np.uint8)
        if value is not None:
            packed = self._codec.to_packed(key, value)
            if len(packed) > 48:
                raise ValueError("snapshot records carry at most 48 program bytes")
            rec[16:16 + len(packed)] = np.frombuffer(packed, dtype=np.uint8)
        self._pending.append(rec)
        self._unflushed[key] = value

    def put(self, key: str, value: Optional[bytes]) -> None:
        if not key:      # the producer's flush record: empty key, empty value (KafkaProducerActorImpl.scala:321-329)
            return
        with self._lock:
            if self._codec is not None:
                self._state_record(key, value)
            else:
                self._overlay[key] = value

    def putIfAbsent(self, key: str, value: bytes) -> Optional[bytes]:  # noqa: N802
        with self._lock:
            cur = self.get(key)
            if cur is None:
                self.put(key, value)
            return cur

    def putAll(self, entries: Sequence[Tuple[str, Optional[bytes]]]) -> None:  # noqa: N802
        for k, v in entries:
            self.put(k, v)

    def delete(self, key: str) -> Optional[bytes]:
        with self._lock:
            cur = self.get(key)
            self.put(key, None)
            return cur

    def get(self, key: str) -> Optional[bytes]:
        if not self._open:
            raise InvalidStateStoreException(N.SGR_ERR_STATE, f"store {self._name} is not open")
        if key in self._overlay:
            return self._overlay[key]
        if key in self._unflushed:            # read-your-writes between put() and flush()
            return self._unflushed[key]
        if not self._folded:
            raise InvalidStateStoreException(N.SGR_ERR_STATE, f"store {self._name} has not been restored yet")
        b = self._engine.get(key)
        if b is None:
            return None
        if self._codec is not None:
            return self._codec.from_packed(key, b)
        return self._formatter(key, b) if self._formatter else b

    def all(self) -> Iterator[Tuple[str, bytes]]:
        with self._lock:
            # KeyValueStore[Bytes, _] iterates in Bytes order: unsigned lexicographic over the UTF-8 key bytes
            keys = sorted(set(self._ingest.keys() if self._ingest is not None else self._keys) | set(self._overlay) | set(self._unflushed),
                          key=lambda k: k.encode("utf-8"))
        for k in keys:
            v = self.get(k)
            if v is not None:
                yield k, v

    def range(self, frm: str, to: str) -> Iterator[Tuple[str, bytes]]:
        lo, hi = frm.encode("utf-8"), to.encode("utf-8")
        for k, v in self.all():
            if lo <= k.encode("utf-8") <= hi:
                yield k, v

    def restoreAll(self, records: Iterable[Tuple[Optional[str], bytes]]) -> None:  # noqa: N802
        """BatchingStateRestoreCallback.restoreAll(Collection[KeyValue[bytes, bytes]]): Kafka Streams hands the restore
        consumer's polls over in batches; one batch = one GPU fold."""
        self.restore(records)

    def approximateNumEntries(self) -> int:  # noqa: N802
        return sum(1 for _ in self.all())

    @property
    def engine(self) -> ReplayEngine:
        return self._engine


class AggregateStateStore:
    """The narrow seam PersistentActor consumes (AggregateStateStoreKafkaStreams.scala:83-89):
    getAggregateBytes(aggregateId): Future[Option[Array[Byte]]] served from a 32-thread pool (ThreadPools.scala:9-11)."""

    def __init__(self, store: GpuReplayKeyValueStore, threads: int = 32):
        self._store = store
        self._pool = ThreadPoolExecutor(max_workers=threads, thread_name_prefix="surge-io")

    def getAggregateBytes(self, aggregateId: str) -> "Future[Optional[bytes]]":  # noqa: N802,N803
        return self._pool.submit(self._store.get, aggregateId)

    def healthCheck(self) -> dict:  # noqa: N802
        return {"name": "aggregate-state-store", "status": "up" if self._store.isOpen() else "down"}

    def stop(self) -> None:
        self._pool.shutdown(wait=True)
