// EventsTopicRebuilder.scala — feeds a GpuReplayKeyValueStore from the EVENTS topic (feed (ii) of GpuReplayPersistencePlugin.scala).
//
// Surge never replays the events topic (SURVEY §0.2): recovery is the KTable restore of the compacted state topic. This component
// is the rebuild that becomes possible with a GPU fold: a plain consumer of the events topic, shaped like the reference's own
// wrapper (modules/common/src/main/scala/surge/kafka/KafkaConsumer.scala:48-105: one poll thread, String key / Array[Byte]
// value, manual commit), that hands every record to store.putEvent, folds once per poll (store.flush) and then commits the
// offsets it has folded for ITS OWN consumer group.
//
//   * isolation.level = read_committed, like the streams consumer (SurgeStateStoreConsumer.scala:38): events of aborted
//     transactions are never folded.
//   * flush records (empty key, KafkaProducerActorImpl.scala:321-329) are dropped by putEvent.
//   * group id = "<applicationId>-gpu-rebuild", NOT the streams application id: the producer's lag gate reads the offsets the
//     streams application id committed on the STATE topic (KafkaProducerActor.scala:54, KafkaProducerActorImpl.scala:684-708);
//     those keep being committed by Kafka Streams after store.flush() (see the plugin's header). This group only records how far
//     the events topic has been folded, so a restarted rebuild resumes instead of starting over; `lag()` exposes it the way
//     KafkaAdminClient.consumerLag does (KafkaAdminClient.scala:44-56).
//   * offsets are committed only AFTER the fold of the poll returned: committed implies readable through get().
//
// Lifecycle and health follow AggregateStateStoreKafkaStreams (AggregateStateStoreKafkaStreams.scala:87-89,127-177): the component
// is a HealthyComponent with a Controllable, registers with the health signal bus on start and unregisters on stop.
//
// NOT COMPILED HERE (no JDK / sbt / jars in the build image).
package surge.gpu

import java.time.Duration
import java.util.Properties
import java.util.concurrent.atomic.{ AtomicBoolean, AtomicLong }
import java.util.regex.Pattern

import org.apache.kafka.clients.consumer.{ ConsumerConfig, KafkaConsumer, OffsetAndMetadata }
import org.apache.kafka.common.TopicPartition
import org.apache.kafka.common.serialization.{ ByteArrayDeserializer, StringDeserializer }
import org.slf4j.LoggerFactory
import surge.core.{ Ack, Controllable }
import surge.health.HealthSignalBusTrait
import surge.internal.health.{ HealthCheck, HealthCheckStatus, HealthyComponent }
import surge.kafka.KafkaTopic

import scala.concurrent.{ ExecutionContext, Future }
import scala.jdk.CollectionConverters._
import scala.util.{ Failure, Success }

class EventsTopicRebuilder(
    eventsTopic: KafkaTopic,
    partitions: Seq[Int], // the partitions this node owns (PartitionAssignments, modules/common/src/main/scala/surge/kafka/PartitionAssignments.scala)
    store: GpuReplayKeyValueStore,
    applicationId: String,
    brokers: Seq[String],
    signalBus: HealthSignalBusTrait,
    extraConsumerProps: Map[String, String] = Map.empty)(implicit ec: ExecutionContext)
    extends HealthyComponent {

  private val log = LoggerFactory.getLogger(getClass)
  private val groupId = s"$applicationId-gpu-rebuild"
  private val componentName = s"gpu-events-rebuilder-${eventsTopic.name}"
  private val running = new AtomicBoolean(false)
  private val failed = new AtomicBoolean(false)
  private val recordsFolded = new AtomicLong(0L)
  @volatile private var thread: Thread = _
  @volatile private var lastLag: Map[TopicPartition, Long] = Map.empty

  private def consumerProps: Properties = {
    val p = new Properties()
    p.put(ConsumerConfig.BOOTSTRAP_SERVERS_CONFIG, brokers.mkString(","))
    p.put(ConsumerConfig.GROUP_ID_CONFIG, groupId)
    p.put(ConsumerConfig.ENABLE_AUTO_COMMIT_CONFIG, "false") // KafkaConsumerHelper.consumerPropsFromConfig, KafkaConsumer.scala:33-35
    p.put(ConsumerConfig.ISOLATION_LEVEL_CONFIG, "read_committed")
    p.put(ConsumerConfig.AUTO_OFFSET_RESET_CONFIG, "earliest") // a rebuild starts at the beginning of the log
    p.put(ConsumerConfig.KEY_DESERIALIZER_CLASS_CONFIG, classOf[StringDeserializer].getName)
    p.put(ConsumerConfig.VALUE_DESERIALIZER_CLASS_CONFIG, classOf[ByteArrayDeserializer].getName)
    extraConsumerProps.foreach { case (k, v) => p.put(k, v) }
    p
  }

  private def pollLoop(): Unit = {
    val consumer = new KafkaConsumer[String, Array[Byte]](consumerProps)
    val tps = partitions.map(p => new TopicPartition(eventsTopic.name, p))
    try {
      consumer.assign(tps.asJava) // explicit assignment: ownership is Surge's partition assignment, not a group rebalance
      while (running.get()) {
        val records = consumer.poll(Duration.ofMillis(200))
        if (!records.isEmpty) {
          val next = scala.collection.mutable.Map.empty[TopicPartition, Long]
          val it = records.iterator()
          while (it.hasNext) {
            val r = it.next()
            store.putEvent(r.key(), r.value()) // null / empty keys (flush records) are dropped there
            next(new TopicPartition(r.topic(), r.partition())) = r.offset() + 1
          }
          store.flush() // ONE GPU fold per poll
          consumer.commitSync(next.map { case (tp, o) => tp -> new OffsetAndMetadata(o) }.asJava) // only what is folded
          recordsFolded.addAndGet(records.count().toLong)
        }
        val end = consumer.endOffsets(tps.asJava).asScala
        lastLag = tps.map(tp => tp -> math.max(0L, end(tp).longValue() - consumer.position(tp))).toMap
      }
    } catch {
      case t: Throwable =>
        // a corrupt record or a CUDA failure must not be retried blindly: report, stop, let supervision decide
        failed.set(true); running.set(false)
        log.error(s"$componentName stopped: ${t.getMessage}", t)
        signalBus.signalWithError(name = "gpu.rebuild.fatal.error", error = surge.health.domain.Error("events-topic rebuild failed", Some(t))).emit()
    } finally consumer.close()
  }

  /** records of the events topic not folded yet, per owned partition (0 everywhere = the table holds the whole log) */
  def lag(): Map[TopicPartition, Long] = lastLag
  def foldedRecords: Long = recordsFolded.get()

  override def healthCheck(): Future[HealthCheck] = Future.successful(
    HealthCheck(
      name = componentName,
      id = groupId,
      status = if (!failed.get() && store.isOpen) HealthCheckStatus.UP else HealthCheckStatus.DOWN,
      details = Some(Map("running" -> running.get().toString, "recordsFolded" -> recordsFolded.get().toString, "lag" -> lastLag.values.sum.toString))))

  override def restartSignalPatterns(): Seq[Pattern] = Seq(Pattern.compile("gpu.rebuild.fatal.error"))

  override val controllable: Controllable = new Controllable {
    override def start(): Future[Ack] = Future {
      if (running.compareAndSet(false, true)) {
        failed.set(false)
        thread = new Thread(() => pollLoop(), componentName)
        thread.setDaemon(true)
        thread.start()
      }
      Ack
    }.andThen {
      case Success(_) =>
        signalBus.register(control = this, componentName = componentName, shutdownSignalPatterns = shutdownSignalPatterns(), restartSignalPatterns = restartSignalPatterns())
      case Failure(e) => log.error(s"unable to start $componentName", e)
    }

    override def stop(): Future[Ack] = Future {
      running.set(false)
      val t = thread
      if (t != null) t.join(10000L)
      Ack
    }.andThen { case _ => signalBus.unregister(control = this, componentName = componentName) }

    override def restart(): Future[Ack] = stop().flatMap(_ => start())
    override def shutdown(): Future[Ack] = stop()
  }
}
