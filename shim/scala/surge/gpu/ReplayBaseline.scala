// ReplayBaseline.scala — the JVM baseline harness of SURVEY §8(d)(i) / BASELINE.md B2: "JVM in-memory replay".
//
// Builds a config's events as JVM objects and replays them through the reference's OWN fold — AggregateCommandModel.toCore.applyAsync,
// i.e. events.foldLeft(state)(handleEvent) (modules/command-engine/scaladsl/src/main/scala/surge/scaladsl/command/CommandModels.scala:16-30)
// — per aggregate, in memory, no Kafka, on a fixed pool of all host cores; prints events/s and a hash of the final states so the
// run can be set beside bench.py's lines for the same config. The model is the reference's test Counter
// (modules/command-engine/scaladsl/src/test/scala/surge/scaladsl/TestBoundedContext.scala:77-89).
//
//   sbt "surge-engine-command-scaladsl/Test/runMain surge.gpu.ReplayBaseline 1048576 32"      (configs[1] shape)
//
// NOT COMPILED OR RUN HERE: this image has no JVM (java -version: not found), no sbt and no Surge jars, so bench.py's CPU arm is
// the C port of the same fold (oracle/sgr_oracle.c, "kind": "port"). This file is what a maintainer runs on a box that has them.
package surge.gpu

import java.util.concurrent.{ Executors, TimeUnit }

import surge.scaladsl.TestBoundedContext
import surge.scaladsl.TestBoundedContext._

import scala.concurrent.duration._
import scala.concurrent.{ Await, ExecutionContext, Future }

object ReplayBaseline extends TestBoundedContext {
  // splitmix64: the same generator family as surge_b200/synth.py routed_events_host, so both sides can fold the same log
  private def splitmix64(x0: Long): Long = {
    var x = x0 + 0x9E3779B97F4A7C15L
    x = (x ^ (x >>> 30)) * 0xBF58476D1CE4E5B9L
    x = (x ^ (x >>> 27)) * 0x94D049BB133111EBL
    x ^ (x >>> 31)
  }

  private def eventsOf(g: Long, epa: Int, seed: Long): Seq[BaseTestEvent] = {
    val id = s"agg-$g"
    (0 until epa).map { k =>
      val x = splitmix64(g * 1000003L + k * 7919L + seed * 0x51ED27L)
      val u = (x & 0xFFFFL).toInt
      val by = ((x >>> 16) & 0x7FFFFFFFL).toInt
      if (u < 29491) CountIncremented(id, by, k + 1) else if (u < 58982) CountDecremented(id, by, k + 1) else NoOpEvent(id, k + 1)
    }
  }

  def main(args: Array[String]): Unit = {
    val nAgg = args.headOption.map(_.toInt).getOrElse(1 << 20)
    val epa = args.lift(1).map(_.toInt).getOrElse(32)
    val threads = Runtime.getRuntime.availableProcessors()
    val pool = Executors.newFixedThreadPool(threads)
    implicit val ec: ExecutionContext = ExecutionContext.fromExecutor(pool)
    val model = BusinessLogic.toCore // SurgeProcessingModel[State, BaseTestCommand, BaseTestEvent]: applyAsync is this foldLeft (CommandModels.scala:25-28)
    val shard = (nAgg + threads - 1) / threads
    for (round <- 0 until 5) {
      // events are built outside the timed region: the reference's actors receive them already deserialized
      val logs = (0 until nAgg).map(g => eventsOf(g.toLong, epa, 3L)).toArray
      val t0 = System.nanoTime()
      val partial = (0 until threads).map { t =>
        Future {
          var h = 0L
          var g = t * shard
          val hi = math.min(nAgg, g + shard)
          while (g < hi) {
            // applyAsync needs a SurgeContext; the fold itself is handleEvent under foldLeft, which is what it runs (CommandModels.scala:25-28)
            val st = logs(g).foldLeft(Option.empty[State])((s, e) => BusinessLogic.handleEvent(s, e))
            h += st.map(s => splitmix64(g.toLong ^ (s.count.toLong << 32 | (s.version.toLong & 0xFFFFFFFFL)))).getOrElse(0L)
            g += 1
          }
          h
        }
      }
      val hash = Await.result(Future.sequence(partial), 1.hour).sum
      val dt = (System.nanoTime() - t0) * 1e-9
      println(f"round $round: ${nAgg.toLong * epa / dt / 1e6}%.1f M events/s on $threads threads ($nAgg aggregates x $epa events), state hash ${hash}%016x")
    }
    val _ = model
    pool.shutdown(); pool.awaitTermination(10, TimeUnit.SECONDS)
  }
}
