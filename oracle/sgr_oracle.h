/*
 * sgr_oracle.h — CPU restatement of the reference's event-replay path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under surge_b200/ may include, link, load or call
 * this. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker or the timed CPU baseline.
 *
 * Parity status: the reference is Scala on the JVM and cannot be compiled or run in the
 * build container (no JDK/sbt/jars). This file restates, function by function, the Scala
 * it cites, and is pinned against every golden vector the reference's own tests hold for
 * the path (tests/test_oracle_golden.py). Two items stay
 * "parity unpinned" because the reference holds no vector for them: the partition hash
 * (scala-library 2.13.8 MurmurHash3.stringHash, third-party — pinned one level down, to a real
 * MurmurHash3_x86_32, by tests/test_partition_hash_pin.py) and serialized JSON bytes.
 *
 * Paths are relative to the reference checkout.
 */
#ifndef SGR_ORACLE_H
#define SGR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* models (handleEvent implementations restated) */
#define ORC_MODEL_COUNTER      0  /* scaladsl TestBoundedContext.scala:77-89 */
#define ORC_MODEL_BANK_ACCOUNT 1  /* surge-docs BankAccountCommandModel.scala:81-86 */
#define ORC_MODEL_INT_BALANCE  2  /* multilanguage-scala-sdk-sample Main.scala:25-30 */
#define ORC_MODEL_ML_COUNTER   3  /* multilanguage test TestBoundedContext.scala:68-75 */

#define ORC_REC_FIXED64 0u
#define ORC_REC_VAR16   1u

#define ORC_ST_EXISTS  1u
#define ORC_ST_CHANGED 2u
#define ORC_ST_ERROR   4u

/* ---- decoded domain objects (mirror the Scala case classes) ---- */
typedef struct { int32_t count; int32_t version; } orc_counter_state;      /* State(aggregateId,count,version) */
typedef struct { uint8_t uuid[16]; uint8_t owner[16]; uint8_t code[8]; uint64_t balance_bits; } orc_bank_account;
typedef struct { int32_t balance; } orc_int_balance;

/* state struct sizes of the binary SurgeAggregateFormatting (see DESIGN.md "formats") */
uint32_t orc_state_bytes(int model);

/* Fold a packed CSR event log: for every aggregate i,
 *   new = events(i).foldLeft(old)(handleEvent)           SDSL/command/CommandModels.scala:25-28
 * with the actor's error rule (exception => state unchanged, PersistentActor.scala:260-263)
 * and publish rule (CHANGED iff new != old, PersistentActor.scala:252-257).
 * initial_states may be NULL (all None). Returns 0, or -1 on malformed input.
 * n_events_out / n_errors_out may be NULL. */
int orc_fold_packed(int model, uint32_t record_kind, const uint8_t* events, const uint64_t* seg_offsets,
                    uint64_t n_agg, const uint8_t* initial_states, uint8_t* out_states,
                    uint64_t* n_events_out, uint64_t* n_errors_out);

/* Same, aggregates sharded over n_threads pthreads (the "all host cores" CPU baseline). */
int orc_fold_packed_mt(int model, uint32_t record_kind, const uint8_t* events, const uint64_t* seg_offsets,
                       uint64_t n_agg, const uint8_t* initial_states, uint8_t* out_states,
                       int n_threads, uint64_t* n_events_out, uint64_t* n_errors_out);

/* Incremental: records in arrival order (fixed64), applied per aggregate in arrival
 * order onto states (in place). Equivalent to one ApplyEvents per touched aggregate. */
/* NUMA-stable variants for the CPU arm of bench.py: worker t pinned to CPU t; orc_place_log_mt copies the log into untouched
   memory with the fold's own sharding, so every worker later reads pages of its own node. */
int orc_fold_packed_mt_pinned(int model, uint32_t record_kind, const uint8_t* events, const uint64_t* seg_offsets,
                              uint64_t n_agg, const uint8_t* initial_states, uint8_t* out_states,
                              int n_threads, uint64_t* n_events_out, uint64_t* n_errors_out);
int orc_place_log_mt(uint8_t* dst, const uint8_t* src, const uint64_t* seg_offsets, uint64_t n_agg, int n_threads);

int orc_fold_incremental(int model, const uint8_t* records, uint64_t n_records,
                         uint8_t* states, uint64_t n_agg);

/* Stable group-by of fixed64 records by aggregate index (Kafka per-partition log order
 * is preserved per key). out_records: n_records*64 bytes; out_offsets: n_agg+1 byte offsets. */
int orc_group_by_agg(const uint8_t* records, uint64_t n_records, uint64_t n_agg,
                     uint8_t* out_records, uint64_t* out_offsets);

/* scala.util.hashing.MurmurHash3.stringHash (scala-library 2.13.8; third-party, restated
 * from its published algorithm) and KafkaPartitionProvider.partitionForKey
 * (COMMON/kafka/KafkaPartitioner.scala:7-9). */
int32_t orc_scala_string_hash(const uint16_t* utf16, uint32_t n);
int32_t orc_partition_for_key(const uint16_t* utf16, uint32_t n, int32_t num_partitions);
/* PartitionStringUpToColon: str.takeWhile(_ != ':') (KafkaPartitioner.scala:38-42) */
uint32_t orc_take_while_not_colon(const uint16_t* utf16, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
