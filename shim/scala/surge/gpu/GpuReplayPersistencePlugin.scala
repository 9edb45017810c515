// GpuReplayPersistencePlugin.scala — the config-selected drop-in behind Surge's state-store seam.
//
//   surge.kafka-streams.state-store-plugin = "gpu-replay"
//   gpu-replay.plugin-class = "surge.gpu.GpuReplayPersistencePlugin"
//
// Implements, without touching SurgeCommand / AggregateRef / SurgeModel:
//   trait SurgeKafkaStreamsPersistencePlugin { def createSupplier(storeName: String): KeyValueBytesStoreSupplier; def enableLogging: Boolean }
//     modules/common/src/main/scala/surge/kafka/streams/SurgeKafkaStreamsPersistencePlugin.scala:12-15
// and returns a KeyValueStore[Bytes, Array[Byte]] shaped like the in-tree example
//     modules/common/src/test/scala/surge/kafka/streams/SingleExceptionThrowingKeyValueStore.scala:18-91.
//
// HOW THE STORE IS FED (the data flow of INTEGRATION.md §1; round-1 review: the previous version parked state records in a JVM
// map and waited for a changelog restore that withLoggingDisabled() guarantees never fires):
//
//   (i)  STATE TOPIC — what Kafka Streams does with this store today. The topology is builder.table(stateTopic, materialized)
//        (SurgeStateStoreConsumer.scala:57-76): the stream thread calls put(aggregateId, serializedState) for every record of
//        the compacted state topic, null = delete (SurgeModel.scala:62-64), then flush() before it commits the offsets of the
//        streams application id. Here put() turns the record into one SNAPSHOT event (or TOMBSTONE) of the registered fold
//        program and flush() folds the batch on the GPU: last write wins per key, exactly the KTable. Because the fold happens
//        inside flush(), the offsets Kafka Streams commits afterwards cover only state that get() can already serve — the lag
//        gate of KafkaProducerActorImpl.scala:684-708 (KafkaAdminClient.consumerLag, KafkaAdminClient.scala:44-56) keeps its
//        meaning with no extra code (SURVEY §8 f2).
//   (ii) EVENTS TOPIC — the rebuild Surge cannot do today (SURVEY §0.2). EventsTopicRebuilder (same package) consumes the events
//        topic with a plain read_committed consumer and calls putEvent(); the same flush() folds events and snapshots in
//        arrival order onto the live table.
//
// The GPU table holds fixed-size packed structs; the topic holds the model's serialized bytes (aggregateWriteFormatting,
// SurgeModel.scala:57-65). A GpuStateCodec registered next to the fold program converts in both directions; a model without one
// cannot use this plugin (sgr_register_program declines it) and keeps the stock RocksDB store.
//
// NOT COMPILED HERE (no JDK / sbt / jars in the build image). The executable twin that the parity tests drive is
// surge_b200/store.py; both are thin adapters over the same C ABI (include/sgr.h).
package surge.gpu

import java.nio.{ ByteBuffer, ByteOrder }
import java.util
import java.util.concurrent.ConcurrentHashMap
import java.util.concurrent.locks.ReentrantReadWriteLock

import org.apache.kafka.common.utils.Bytes
import org.apache.kafka.streams.KeyValue
import org.apache.kafka.streams.errors.InvalidStateStoreException
import org.apache.kafka.streams.processor.{ ProcessorContext, StateStore }
import org.apache.kafka.streams.state.{ KeyValueBytesStoreSupplier, KeyValueIterator, KeyValueStore }
import org.slf4j.LoggerFactory
import surge.kafka.streams.SurgeKafkaStreamsPersistencePlugin

/** serialized state bytes (what the topic and the actors hold) <-> the packed program bytes of the GPU table */
trait GpuStateCodec {

  /** program bytes of the state struct: sgr_fold_program.state_bytes - 8. At most 48 for this shim (one fixed 64-byte snapshot
   *  record carries them at +16); wider states need VAR16 snapshot records (include/sgr.h), not wired here. */
  def programBytes: Int

  /** aggregateReadFormatting.readState, then the packed layout of surge_b200/formats.py for the model */
  def toPacked(aggregateId: String, serialized: Array[Byte]): Array[Byte]

  /** the inverse: aggregateWriteFormatting.writeState of the state the packed bytes stand for */
  def fromPacked(aggregateId: String, packed: Array[Byte]): Array[Byte]
}

/** A model registers the declarative form of its handleEvent once, next to the JVM handler: the fold program (its rules for the
 *  model's event classes PLUS one CREATE+SET rule `snapshotType` that copies program bytes from +16 and one TOMBSTONE rule
 *  `tombstoneType`; surge_b200/dsl.py emits both with `with_snapshot_rules`) and the codec. */
object GpuFoldPrograms {
  final case class Registration(program: ByteBuffer, codec: GpuStateCodec, snapshotType: Int, tombstoneType: Int)
  @volatile private var registration: Option[Registration] = None
  def register(packedSgrFoldProgram: ByteBuffer, codec: GpuStateCodec, snapshotType: Int, tombstoneType: Int): Unit = {
    require(codec.programBytes > 0 && codec.programBytes <= 48 && codec.programBytes % 4 == 0, "snapshot records carry at most 48 program bytes")
    registration = Some(Registration(packedSgrFoldProgram, codec, snapshotType, tombstoneType))
  }
  def current: Registration = registration.getOrElse(throw new IllegalStateException("no GPU fold program registered for this model"))
}

class GpuReplayPersistencePlugin extends SurgeKafkaStreamsPersistencePlugin {
  private val log = LoggerFactory.getLogger(getClass)
  // No changelog: the source topic IS the log of this table (SurgeStateStoreConsumer.scala:63-75 builds the topology
  // withLoggingDisabled() and unoptimised in that case, so every state record reaches put()).
  override def enableLogging: Boolean = false

  override def createSupplier(storeName: String): KeyValueBytesStoreSupplier = {
    // The loader swallows every failure and falls back to RocksDB (SurgeKafkaStreamsPersistencePlugin.scala:34-47):
    // be loud here so a silent fallback is visible in the logs, and fail in the constructor path rather than later.
    GpuFoldPrograms.current
    log.warn(s"GPU replay state store '$storeName' selected; if RocksDB metrics appear the plugin failed to load")
    new KeyValueBytesStoreSupplier {
      override def name(): String = storeName
      override def get(): KeyValueStore[Bytes, Array[Byte]] = new GpuReplayKeyValueStore(storeName)
      override def metricsScope(): String = "gpu-replay"
    }
  }
}

class GpuReplayKeyValueStore(storeName: String) extends KeyValueStore[Bytes, Array[Byte]] {
  private val log = LoggerFactory.getLogger(getClass)
  private val reg = GpuFoldPrograms.current
  private var handle: Long = 0L
  @volatile private var open = false
  // write side: ONE thread (the Kafka Streams stream thread, or the rebuilder's poll thread) — guarded by `lock` against the
  // 32 reader threads of ThreadPools.ioBoundContext that call get()/all()/range()
  private val lock = new ReentrantReadWriteLock()
  private val pending = new java.io.ByteArrayOutputStream()
  private var pendingRecords = 0L
  private val keyIndex = new util.HashMap[String, java.lang.Long]() // aggregate id -> dense slot (first-seen order)
  private val keys = new util.ArrayList[String]()
  // read-your-writes between put() and flush(): the not-yet-folded value of a key (None = deleted)
  private val unflushed = new ConcurrentHashMap[String, Option[Array[Byte]]]()
  private var capacity = 0L
  private var folded = false
  private var loadedKeys = -1

  private def check(rc: Int): Unit = if (rc != Native.OK) {
    val msg = Native.lastError(handle)
    if (rc == Native.ERR_STATE) throw new InvalidStateStoreException(msg) else throw new RuntimeException(s"sgr error $rc: $msg")
  }

  override def name(): String = storeName
  // like the in-memory stores of Kafka Streams: nothing on local disk, the table is rebuilt from the topic after a restart
  override def persistent(): Boolean = false
  override def isOpen: Boolean = open

  override def init(context: ProcessorContext, root: StateStore): Unit = {
    handle = Native.create(0) // throws when there is no usable GPU: no CPU fallback, the stream thread dies loudly
    check(Native.registerProgram(handle, reg.program))
    // The restore callback exists for stores with a changelog (SingleExceptionThrowingKeyValueStore.scala:84-86). This store has
    // none (enableLogging = false), so Kafka Streams never calls it; it is registered because StateStore.init must register the
    // root store, and it does the right thing if a future topology does restore through it.
    context.register(root, (key: Array[Byte], value: Array[Byte]) => put(Bytes.wrap(key), value))
    open = true
  }

  // ---------------------------------------------------------------- write side
  private def slotOf(aggregateId: String): Long = {
    var s = keyIndex.get(aggregateId)
    if (s == null) { s = java.lang.Long.valueOf(keys.size().toLong); keyIndex.put(aggregateId, s); keys.add(aggregateId) }
    s.longValue()
  }

  private def appendRecord(eventType: Int, seq: Int, slot: Long, payload: Array[Byte]): Unit = {
    val rec = ByteBuffer.allocate(64).order(ByteOrder.LITTLE_ENDIAN)
    rec.putInt(0, eventType); rec.putInt(4, seq); rec.putLong(8, slot)
    if (payload != null) { rec.position(16); rec.put(payload, 0, math.min(payload.length, 48)) }
    pending.write(rec.array()); pendingRecords += 1
  }

  /** (i) one record of the STATE topic: KTable semantics, last write wins, null deletes (SurgeStateStoreConsumer.scala:57-76) */
  override def put(key: Bytes, value: Array[Byte]): Unit = {
    val id = new String(key.get(), "UTF-8")
    if (id.isEmpty) return // the producer's flush record: empty key, empty value (KafkaProducerActorImpl.scala:321-329)
    lock.writeLock().lock()
    try {
      if (value == null) appendRecord(reg.tombstoneType, 0, slotOf(id), null)
      else appendRecord(reg.snapshotType, 0, slotOf(id), reg.codec.toPacked(id, value))
      unflushed.put(id, Option(value))
    } finally lock.writeLock().unlock()
  }

  /** (ii) one record of the EVENTS topic: value = the model's packed event (u32 type, u32 seq, payload; formats.py) */
  def putEvent(recordKey: String, packedEvent: Array[Byte]): Unit = if (recordKey != null && recordKey.nonEmpty) {
    require(packedEvent.length >= 8 && packedEvent.length <= 56, "packed event: u32 type, u32 seq, up to 48 payload bytes")
    val id = recordKey.takeWhile(_ != ':') // PartitionStringUpToColon, KafkaPartitioner.scala:38-42
    val b = ByteBuffer.wrap(packedEvent).order(ByteOrder.LITTLE_ENDIAN)
    lock.writeLock().lock()
    try {
      appendRecord(b.getInt(0), b.getInt(4), slotOf(id), util.Arrays.copyOfRange(packedEvent, 8, packedEvent.length))
      unflushed.remove(id) // an event supersedes an unflushed snapshot view: readers see the folded table again after flush()
    } finally lock.writeLock().unlock()
  }

  override def putIfAbsent(key: Bytes, value: Array[Byte]): Array[Byte] = { val cur = get(key); if (cur == null) put(key, value); cur }
  override def putAll(entries: util.List[KeyValue[Bytes, Array[Byte]]]): Unit = entries.forEach(kv => put(kv.key, kv.value))
  override def delete(key: Bytes): Array[Byte] = { val cur = get(key); put(key, null); cur }

  /** Kafka Streams calls this before it commits offsets: everything put() so far is folded when it returns. */
  override def flush(): Unit = {
    lock.writeLock().lock()
    try {
      if (pendingRecords == 0L && folded) return
      val batch = pending.toByteArray; pending.reset()
      val n = pendingRecords; pendingRecords = 0L
      if (!folded || keys.size() > capacity) growTable()
      if (n > 0) {
        val direct = ByteBuffer.allocateDirect(batch.length); direct.put(batch); direct.flip()
        check(Native.foldIncremental(handle, direct, n))
      }
      if (keys.size() != loadedKeys) loadKeyTable() // new ids inside the current capacity
      unflushed.clear()
    } finally lock.writeLock().unlock()
  }

  /** Make room for the keys seen so far: the table is resized on the device, content kept, new slots None (sgr_grow_states). */
  private def growTable(): Unit = {
    val newCapacity = math.max(2L * keys.size(), 1024L)
    check(Native.growStates(handle, newCapacity))
    capacity = newCapacity
    folded = true
    loadKeyTable()
  }

  /** key table for sgr_get: ids in slot order, unused slots get unreachable placeholder keys */
  private def loadKeyTable(): Unit = {
    val blob = new java.io.ByteArrayOutputStream()
    val offs = ByteBuffer.allocateDirect((capacity.toInt + 1) * 4).order(ByteOrder.LITTLE_ENDIAN)
    offs.putInt(0)
    var i = 0
    while (i < capacity) {
      val bytes = (if (i < keys.size()) keys.get(i) else " unused-" + i).getBytes("UTF-8")
      blob.write(bytes); offs.putInt(blob.size()); i += 1
    }
    val kb = ByteBuffer.allocateDirect(math.max(blob.size(), 1)); kb.put(blob.toByteArray); kb.flip(); offs.flip()
    check(Native.loadKeys(handle, kb, offs, capacity))
    loadedKeys = keys.size()
  }

  // ---------------------------------------------------------------- read side
  /** The recovery read: AggregateStateStoreKafkaStreams.getAggregateBytes ends here (KafkaStreamsKeyValueStore.scala:24-26).
   *  A fresh array every time, as RocksDB returns (SURVEY §8b ownership). */
  override def get(key: Bytes): Array[Byte] = {
    if (!open) throw new InvalidStateStoreException(s"store $storeName is not open")
    val id = new String(key.get(), "UTF-8")
    val u = unflushed.get(id)
    if (u != null) return u.map(_.clone()).orNull
    if (!folded) return null // nothing was ever put: an empty table, not an error (KTable miss)
    val packed = Native.get(handle, key.get()) // null == None; InvalidStateStoreException while the table is being rebuilt
    if (packed == null) null else reg.codec.fromPacked(id, packed)
  }

  /** Snapshot of the keys in Bytes order (unsigned lexicographic over the UTF-8 bytes), values resolved lazily through get(). */
  private def orderedIterator(from: Bytes, to: Bytes): KeyValueIterator[Bytes, Array[Byte]] = {
    if (!open) throw new InvalidStateStoreException(s"store $storeName is not open")
    lock.readLock().lock()
    val sorted: Array[Bytes] =
      try {
        val all = new util.TreeSet[Bytes]() // Bytes.compareTo is the store order of Kafka Streams
        keys.forEach(k => all.add(Bytes.wrap(k.getBytes("UTF-8"))))
        unflushed.keySet().forEach(k => all.add(Bytes.wrap(k.getBytes("UTF-8"))))
        val view = if (from == null && to == null) all else all.subSet(if (from == null) all.first() else from, true, if (to == null) all.last() else to, true)
        view.toArray(new Array[Bytes](0))
      } finally lock.readLock().unlock()
    new KeyValueIterator[Bytes, Array[Byte]] {
      private var i = 0
      private var nextKv: KeyValue[Bytes, Array[Byte]] = _
      private def advance(): Unit = {
        nextKv = null
        while (nextKv == null && i < sorted.length) {
          val v = get(sorted(i)) // deleted / None entries are skipped, like tombstoned RocksDB keys
          if (v != null) nextKv = new KeyValue(sorted(i), v)
          i += 1
        }
      }
      advance()
      override def hasNext: Boolean = nextKv != null
      override def next(): KeyValue[Bytes, Array[Byte]] = {
        if (nextKv == null) throw new java.util.NoSuchElementException
        val r = nextKv; advance(); r
      }
      override def peekNextKey(): Bytes = { if (nextKv == null) throw new java.util.NoSuchElementException; nextKv.key }
      override def close(): Unit = ()
    }
  }

  override def range(from: Bytes, to: Bytes): KeyValueIterator[Bytes, Array[Byte]] = orderedIterator(from, to)
  override def all(): KeyValueIterator[Bytes, Array[Byte]] = orderedIterator(null, null)

  /** upper bound, like RocksDB's estimate: ids ever seen (deleted ones included until the next rebuild) + unflushed new ones */
  override def approximateNumEntries(): Long = {
    lock.readLock().lock()
    try {
      var n = keys.size().toLong
      unflushed.keySet().forEach(k => if (!keyIndex.containsKey(k)) n += 1)
      n
    } finally lock.readLock().unlock()
  }

  override def close(): Unit = {
    lock.writeLock().lock()
    try { open = false; if (handle != 0L) Native.destroy(handle); handle = 0L }
    finally lock.writeLock().unlock()
    log.info(s"GPU replay state store '$storeName' closed")
  }
}
