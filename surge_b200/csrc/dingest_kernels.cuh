// dingest_kernels.cuh — device-side decode of Kafka record batches (launch interface of dingest_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgr {

// error codes a kernel leaves in DgBatch::err (0 = fine); dingest.cu turns them into messages
enum DgErr : uint32_t {
  DG_OK = 0, DG_CRC = 1, DG_LZ4_HEADER = 2, DG_LZ4_BLOCK = 3, DG_LZ4_SEQUENCE = 4, DG_LZ4_CHECKSUM = 5, DG_LZ4_TOO_LARGE = 6,
  DG_RECORD_LENGTH = 7, DG_RECORD_MALFORMED = 8, DG_RECORD_COUNT = 9, DG_VALUE_LENGTH = 10, DG_ID_LENGTH = 11, DG_STRAY_BYTES = 12,
  DG_ARENA_FULL = 13,   // not a data error: the batch's arena claim did not fit (the host lays the arena out exactly and repeats)
};

// One data batch that survived the host's header walk (control batches, aborted transactions and anything below the partition's
// position never reach the device).
struct DgBatch {
  uint64_t src_off;      // byte offset of the batch (its baseOffset field) inside the wire buffer
  int64_t base_offset;   // first offset of the batch
  int64_t min_offset;    // records below this offset were decoded by an earlier call: duplicates, dropped
  uint32_t total_len;    // 12 + batchLength
  uint32_t n_records;    // recordsCount of the header
  uint32_t codec;        // 0 none, 3 lz4
  uint32_t stored_crc;   // CRC-32C field of the header
  uint32_t rec_base;     // index of the batch's first record in the per-record tables
  uint32_t dsize;        // out (size pass): decompressed bytes of the records section (lz4), else its stored length
  uint64_t arena_off;    // in (decode pass): where the decompressed section goes
  uint32_t err;          // out: DgErr
  uint32_t err_record;   // out: record index the error refers to
};

struct DgDict {           // device id dictionary: open addressing on a 64-bit hash, ids compared byte for byte
  unsigned long long* tags;   // [slots] 0 = empty, else the id's hash (never 0)
  uint32_t* slot_idx;         // [slots] dense index + 1 once the owner has published the id (0 = not yet)
  uint2* key_ref;             // [max_keys] (arena offset in 8-byte units, length) of dense index i
  uint8_t* arena;             // id bytes, 8-byte aligned entries
  unsigned long long* ctl;    // [0] n_keys [1] arena bytes used [2] records dropped as markers [3] null values [4] duplicates
                              // [5] dictionary overflow (keys or arena) [6] packed records written (non-holes)
                              // [8] [9] [10] decode arena: bytes claimed, capacity, overflow flag (dg_launch_crc_size_fast)
  uint64_t slots_mask;        // slots - 1 (power of two)
  uint64_t max_keys, arena_cap;
};

struct DgParse {
  const uint8_t* wire;
  const uint8_t* arena;
  DgBatch* batches;
  uint32_t n_batches;
  uint32_t rec_begin;            // first record slot this launch parses
  uint32_t n_records;            // one past the last record slot this launch parses
  const uint32_t* rec_off;       // [n_records] offset of the record (its length varint) inside the batch's records section
  const uint32_t* rec_batch;     // [n_records] batch of the record
  uint8_t* out;                  // [n_records] packed 64-byte records; dropped records become holes (agg == ~0)
  int32_t null_value_type;       // -1: keyed records with a null value are dropped; else they become events of this type
  DgDict dict;
};

cudaError_t dg_launch_crc_size(const uint8_t* wire, DgBatch* batches, uint32_t n, cudaStream_t st);
// batches: the sub-array to process (n of them), whose first element has index `index_base` in the full array (what rec_batch records)
cudaError_t dg_launch_decode_walk(const uint8_t* wire, uint8_t* arena, DgBatch* batches, uint32_t n, uint32_t index_base, uint32_t* rec_off, uint32_t* rec_batch, cudaStream_t st);
cudaError_t dg_launch_parse(const DgParse& p, cudaStream_t st);
// second generation (register-window decode, device-side arena claims); wire and arena need 64 readable bytes past their content
// nbytes rounded up to 16: both buffers need that much room; host_mapped is the DEVICE address of page-locked, mapped host memory
cudaError_t dg_copy_from_mapped_host(const void* host_mapped, void* dst, uint64_t nbytes, cudaStream_t st);
cudaError_t dg_prepare();   // uploads the CRC tables (a synchronous copy: call it before anything runs on other streams)
cudaError_t dg_launch_crc_size_fast(const uint8_t* wire, DgBatch* batches, uint32_t n, unsigned long long* arena_ctl, cudaStream_t st);
// arena_ctl as given to dg_launch_crc_size_fast for the same batches (nullptr: dsize is exact, not a slot capacity)
cudaError_t dg_launch_decode_walk_fast(const uint8_t* wire, uint8_t* arena, DgBatch* batches, uint32_t n, uint32_t index_base, uint32_t* rec_off, uint32_t* rec_batch,
                                       unsigned long long* arena_ctl, cudaStream_t st);
cudaError_t dg_gather_keys(const DgDict& d, uint64_t from, uint32_t n, uint32_t* d_offs, uint8_t* d_bytes, uint32_t* d_tmp, cudaStream_t st);
uint32_t dg_crc32c_host_reference_polynomial();   // 0x82F63B78: the tables of the device CRC are built from it at first use

}  // namespace sgr
