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
// NOT COMPILED HERE (no JDK / sbt / jars in the build image). The executable twin that the parity tests drive is
// surge_b200/store.py; both are thin adapters over the same C ABI (include/sgr.h).
package surge.gpu

import java.nio.{ ByteBuffer, ByteOrder }
import java.util

import org.apache.kafka.common.utils.Bytes
import org.apache.kafka.streams.KeyValue
import org.apache.kafka.streams.errors.InvalidStateStoreException
import org.apache.kafka.streams.processor.{ ProcessorContext, StateStore }
import org.apache.kafka.streams.state.{ KeyValueBytesStoreSupplier, KeyValueIterator, KeyValueStore }
import org.slf4j.LoggerFactory
import surge.kafka.streams.SurgeKafkaStreamsPersistencePlugin

/** A model registers the declarative form of its handleEvent once, next to the JVM handler. */
object GpuFoldPrograms {
  @volatile private var program: Option[ByteBuffer] = None
  def register(packedSgrFoldProgram: ByteBuffer): Unit = program = Some(packedSgrFoldProgram)
  def current: ByteBuffer = program.getOrElse(throw new IllegalStateException("no GPU fold program registered for this model"))
}

class GpuReplayPersistencePlugin extends SurgeKafkaStreamsPersistencePlugin {
  private val log = LoggerFactory.getLogger(getClass)
  // Kafka Streams must not restore this store from a changelog: it rebuilds from the events topic
  // (SurgeStateStoreConsumer.scala:63-75 builds the topology withLoggingDisabled() in that case).
  override def enableLogging: Boolean = false

  override def createSupplier(storeName: String): KeyValueBytesStoreSupplier = {
    // The loader swallows every failure and falls back to RocksDB (SurgeKafkaStreamsPersistencePlugin.scala:34-47):
    // be loud here so a silent fallback is visible in the logs.
    log.warn(s"GPU replay state store '$storeName' selected; if RocksDB metrics appear the plugin failed to load")
    new KeyValueBytesStoreSupplier {
      override def name(): String = storeName
      override def get(): KeyValueStore[Bytes, Array[Byte]] = new GpuReplayKeyValueStore(storeName)
      override def metricsScope(): String = "gpu-replay"
    }
  }
}

class GpuReplayKeyValueStore(storeName: String) extends KeyValueStore[Bytes, Array[Byte]] {
  private var handle: Long = 0L
  private var open = false
  private val pending = new java.io.ByteArrayOutputStream()
  private val keyIndex = new util.HashMap[String, java.lang.Long]()
  private val overlay = new util.concurrent.ConcurrentHashMap[String, Option[Array[Byte]]]()
  private var capacity = 0L
  private var folded = false
  private var loadedKeys = -1

  private def check(rc: Int): Unit = if (rc != Native.OK) {
    val msg = Native.lastError(handle)
    if (rc == Native.ERR_STATE) throw new InvalidStateStoreException(msg) else throw new RuntimeException(s"sgr error $rc: $msg")
  }

  override def name(): String = storeName
  override def persistent(): Boolean = false
  override def isOpen: Boolean = open

  override def init(context: ProcessorContext, root: StateStore): Unit = {
    handle = Native.create(0)
    check(Native.registerProgram(handle, GpuFoldPrograms.current))
    // restore hook, as in SingleExceptionThrowingKeyValueStore.scala:84-86
    context.register(root, (key: Array[Byte], value: Array[Byte]) => putEvent(new String(key, "UTF-8"), value))
    open = true
  }

  /** One record of the events topic. Empty-key flush markers (KafkaProducerActorImpl.scala:321-329) are dropped. */
  def putEvent(recordKey: String, packedEvent: Array[Byte]): Unit = if (recordKey != null && recordKey.nonEmpty) {
    val id = recordKey.takeWhile(_ != ':') // PartitionStringUpToColon, KafkaPartitioner.scala:38-42
    val slot = keyIndex.computeIfAbsent(id, _ => java.lang.Long.valueOf(keyIndex.size().toLong))
    val rec = ByteBuffer.wrap(packedEvent.clone()).order(ByteOrder.LITTLE_ENDIAN)
    rec.putLong(8, slot)
    pending.write(rec.array())
  }

  override def flush(): Unit = {
    val batch = pending.toByteArray; pending.reset()
    val direct = ByteBuffer.allocateDirect(batch.length); direct.put(batch); direct.flip()
    if (!folded || keyIndex.size() > capacity) growTable()
    check(Native.foldIncremental(handle, direct, batch.length / 64))
    if (keyIndex.size() != loadedKeys) loadKeyTable() // new ids inside the current capacity
  }

  /** Make room for the keys seen so far: the table is resized on the device, content kept, new slots None (sgr_grow_states). */
  private def growTable(): Unit = {
    val newCapacity = math.max(2L * keyIndex.size(), 1024L)
    check(Native.growStates(handle, newCapacity))
    capacity = newCapacity
    folded = true
    loadKeyTable()
  }

  // ---- raw broker bytes (a consumer that hands over fetch responses undecoded): decode natively, fold, report offsets.
  // A store is fed either through putEvent or through restoreRecordBatches, not both (each side keeps its own id dictionary).
  private lazy val ingest: Long = Native.ingestCreate()

  /** bytes of one fetch response for `partition` + its aborted transactions; decoded as a read_committed consumer would
   *  (SurgeStateStoreConsumer.scala:38). Returns the number of event records appended to the pending batch. */
  def restoreRecordBatches(partition: Int, fetch: ByteBuffer, abortedProducerIds: Array[Long], abortedFirstOffsets: Array[Long]): Long = {
    if (abortedProducerIds.nonEmpty) Native.ingestSetAborted(ingest, partition, abortedProducerIds, abortedFirstOffsets)
    Native.ingestRecordBatches(ingest, partition, fetch, fetch.remaining().toLong)
  }

  /** fold what restoreRecordBatches decoded; afterwards committedOffset(partition) is what the consumer acting for this
   *  store commits for the streams application id, so that the producer's lag check reaches zero exactly when get() can
   *  serve the state (KafkaProducerActorImpl.scala:684-708, KafkaAdminClient.scala:44-56). */
  def flushRecordBatches(): Unit = { check(Native.foldIngested(handle, ingest)); folded = true }
  def committedOffset(partition: Int): Long = Native.ingestOffsets(ingest, partition)(1)

  /** key table for sgr_get: ids in slot order, unused slots get unreachable placeholder keys */
  private def loadKeyTable(): Unit = {
    val newCapacity = capacity
    val ids = new Array[String](newCapacity.toInt)
    keyIndex.forEach((id, slot) => ids(slot.intValue()) = id)
    val blob = new java.io.ByteArrayOutputStream()
    val offs = ByteBuffer.allocateDirect((newCapacity.toInt + 1) * 4).order(ByteOrder.LITTLE_ENDIAN)
    offs.putInt(0)
    var i = 0
    while (i < newCapacity) {
      val bytes = (if (ids(i) != null) ids(i) else "\u0000unused-" + i).getBytes("UTF-8")
      blob.write(bytes); offs.putInt(blob.size()); i += 1
    }
    val keys = ByteBuffer.allocateDirect(math.max(blob.size(), 1)); keys.put(blob.toByteArray); keys.flip(); offs.flip()
    check(Native.loadKeys(handle, keys, offs, newCapacity))
    loadedKeys = keyIndex.size()
  }

  // KTable semantics for state records: last write wins, null deletes (SurgeStateStoreConsumer.scala:57-76)
  override def put(key: Bytes, value: Array[Byte]): Unit = overlay.put(key.toString, Option(value))
  override def putIfAbsent(key: Bytes, value: Array[Byte]): Array[Byte] = { val cur = get(key); if (cur == null) put(key, value); cur }
  override def putAll(entries: util.List[KeyValue[Bytes, Array[Byte]]]): Unit = entries.forEach(kv => put(kv.key, kv.value))
  override def delete(key: Bytes): Array[Byte] = { val cur = get(key); overlay.put(key.toString, None); cur }

  /** The recovery read: AggregateStateStoreKafkaStreams.getAggregateBytes ends here (KafkaStreamsKeyValueStore.scala:24-26). */
  override def get(key: Bytes): Array[Byte] = {
    if (!open) throw new InvalidStateStoreException(s"store $storeName is not open")
    Option(overlay.get(key.toString)) match {
      case Some(v) => v.orNull
      case None    => Native.get(handle, key.get()) // null == None; throws InvalidStateStoreException before the first fold
    }
  }

  override def range(from: Bytes, to: Bytes): KeyValueIterator[Bytes, Array[Byte]] = throw new UnsupportedOperationException("next round (f3)")
  override def all(): KeyValueIterator[Bytes, Array[Byte]] = throw new UnsupportedOperationException("next round (f3)")
  override def approximateNumEntries(): Long = keyIndex.size().toLong
  override def close(): Unit = { open = false; if (handle != 0L) Native.destroy(handle); handle = 0L }
}
