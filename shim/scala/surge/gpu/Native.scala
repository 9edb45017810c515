// Native.scala — JNI face of libsgr.so (include/sgr.h). One `external` per C entry point, same order, same meaning.
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK / sbt / Surge jars); shipped as the binding a maintainer adds
// next to modules/common. The C side of these stubs is shim/jni/sgr_jni.c.
package surge.gpu

import java.nio.ByteBuffer

object Native {
  System.loadLibrary("sgr_jni") // links libsgr.so

  // status codes of include/sgr.h
  final val OK = 0
  final val ERR_STATE = -8 // -> org.apache.kafka.streams.errors.InvalidStateStoreException

  @native def create(device: Int): Long                                            // sgr_create
  @native def destroy(handle: Long): Int                                           // sgr_destroy
  @native def lastError(handle: Long): String                                      // sgr_last_error
  @native def registerProgram(handle: Long, program: ByteBuffer): Int              // sgr_register_program (packed sgr_fold_program)
  @native def loadEvents(handle: Long, events: ByteBuffer, nbytes: Long, segOffsets: ByteBuffer, nAgg: Long): Int // sgr_load_events
  @native def loadUnsorted(handle: Long, records: ByteBuffer, nRecords: Long, nAgg: Long): Int // sgr_load_unsorted
  @native def setInitialStates(handle: Long, states: ByteBuffer, nAgg: Long): Int  // sgr_set_initial_states
  @native def fold(handle: Long): Int                                              // sgr_fold
  @native def foldIncremental(handle: Long, records: ByteBuffer, nRecords: Long): Int // sgr_fold_incremental
  @native def loadKeys(handle: Long, keys: ByteBuffer, keyOffsets: ByteBuffer, nAgg: Long): Int // sgr_load_keys
  /** returns null for None, the program bytes otherwise; throws on a non-OK status */
  @native def get(handle: Long, key: Array[Byte]): Array[Byte]                     // sgr_get
  @native def exportStates(handle: Long, out: ByteBuffer, changedBits: ByteBuffer): Int // sgr_export_states
  @native def partitionForKey(key: Array[Byte], numPartitions: Int, upToColon: Boolean): Int // sgr_partition_for_key_utf8

  // raw record batches in, committed offsets out (include/sgr.h "ingest")
  @native def ingestCreate(): Long                                                 // sgr_ingest_create
  @native def ingestDestroy(ingest: Long): Int                                     // sgr_ingest_destroy
  @native def ingestSetValueFraming(ingest: Long, framing: Int): Int               // sgr_ingest_set_value_framing (0 packed, 1 protobuf Event, 2 JSON)
  @native def ingestSetNullValueType(ingest: Long, eventType: Int): Int            // sgr_ingest_set_null_value_type (state-topic tombstones)
  @native def ingestSetAborted(ingest: Long, partition: Int, producerIds: Array[Long], firstOffsets: Array[Long]): Int // sgr_ingest_set_aborted
  /** decodes one fetch response's bytes for `partition`; returns the number of packed records appended; throws on malformed input */
  @native def ingestRecordBatches(ingest: Long, partition: Int, data: ByteBuffer, nbytes: Long): Long // sgr_ingest_record_batches
  @native def foldIngested(handle: Long, ingest: Long): Int                        // sgr_fold_ingested
  @native def growStates(handle: Long, nAgg: Long): Int                            // sgr_grow_states
  /** Array(decodedNext, foldedNext) */
  @native def ingestOffsets(ingest: Long, partition: Int): Array[Long]             // sgr_ingest_offsets

  // the same bytes decoded ON THE DEVICE: only the wire bytes cross PCIe (include/sgr.h "device ingest"); packed values only
  @native def dingestCreate(handle: Long, maxKeys: Long, maxIdBytes: Long): Long   // sgr_dingest_create (maxIdBytes 0 = 32 per id)
  @native def dingestDestroy(dingest: Long): Int                                   // sgr_dingest_destroy
  @native def dingestSetNullValueType(dingest: Long, eventType: Int): Int          // sgr_dingest_set_null_value_type
  @native def dingestSetAborted(dingest: Long, partition: Int, producerIds: Array[Long], firstOffsets: Array[Long]): Int // sgr_dingest_set_aborted
  /** queues one fetch response's bytes (a DIRECT buffer, untouched until dingestFold returns); returns the data batches queued */
  @native def dingestSubmit(dingest: Long, partition: Int, data: ByteBuffer, nbytes: Long): Long // sgr_dingest_submit
  /** CRC, lz4, parse, intern and fold of everything submitted, all or nothing; Array(recordsFolded, newAggregateIds) */
  @native def dingestFold(dingest: Long): Array[Long]                              // sgr_dingest_fold
  /** Array(decodedNext, foldedNext) */
  @native def dingestOffsets(dingest: Long, partition: Int): Array[Long]           // sgr_dingest_offsets
  @native def dingestReset(dingest: Long): Int                                     // sgr_dingest_reset
}
