// dingest.cu — host side of the device record-batch decode (include/sgr.h "device ingest"; kernels: dingest_kernels.cu).
//
// The host keeps exactly what is sequential and tiny: the walk over the 61-byte RecordBatch headers of a fetch (boundaries, a
// trailing partial batch, magic), the read_committed bookkeeping of org.apache.kafka consumers (control batches, the aborted
// transactions a fetch response announces) and the partition positions for the lag gate
// (modules/command-engine/core/src/main/scala/surge/internal/kafka/KafkaProducerActorImpl.scala:684-708). The wire bytes go to
// the device as they are; CRC, lz4, record parsing, id interning and the fold never touch the CPU.
//
// Call sequence per poll:   sgr_dingest_submit(partition, fetch bytes)*  ->  sgr_dingest_fold()
// submit = header walk + one asynchronous H2D copy of the fetch (page-locked source memory makes it a single DMA);
// fold   = crc_size -> (dsize back, arena offsets out) -> decode_walk -> parse + intern -> table growth -> the sort-free fold of
//          the decoded records onto the engine's live table; only then do the partitions' positions advance. All or nothing.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/sgr.h"
#include "devbuf.h"
#include "dingest_kernels.cuh"

using namespace sgr;

namespace {
inline uint16_t be16(const uint8_t* p) { return (uint16_t)((p[0] << 8) | p[1]); }
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint64_t be64(const uint8_t* p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }
constexpr uint64_t kBatchHeader = 61;

struct PartState {
  int64_t decoded_next = 0, folded_next = 0;
  bool seen = false;
  std::vector<std::pair<int64_t, int64_t>> aborted;   // (first_offset, producer_id), ascending, not yet reached
  std::unordered_set<int64_t> aborting;               // producers inside an aborted transaction right now
};

// growable device buffer that keeps its content (the staged fetches of one poll accumulate in it)
struct KeepBuf {
  DevBuf b;
  uint64_t used = 0;
  cudaError_t ensure(uint64_t extra, cudaStream_t st) {
    if (used + extra <= b.cap) return cudaSuccess;
    DevBuf nb;
    uint64_t cap = b.cap ? b.cap : (1ull << 20);
    while (cap < used + extra) cap *= 2;
    cudaError_t e = nb.reserve(cap);
    if (e != cudaSuccess) return e;
    if (used) { e = cudaMemcpyAsync(nb.p, b.p, used, cudaMemcpyDeviceToDevice, st); if (e == cudaSuccess) e = cudaStreamSynchronize(st); }
    if (e != cudaSuccess) { nb.release(); return e; }
    b.release();
    b = nb;
    return cudaSuccess;
  }
};
}  // namespace

struct sgr_dingest {
  sgr_engine* eng = nullptr;
  cudaStream_t stream = nullptr;
  std::string last_error;
  std::map<int32_t, PartState> parts;       // committed view (after the last successful fold)
  std::map<int32_t, PartState> staged;      // view after the submissions of the current poll
  int32_t null_value_type = -1;
  // staged submissions
  KeepBuf wire;
  std::vector<DgBatch> batches;
  uint64_t n_record_slots = 0;
  sgr_ingest_stats poll{};                  // statistics of the current poll (host-side parts)
  sgr_ingest_stats total{};
  // device scratch
  DevBuf d_batches, arena, rec_off, rec_batch, out;
  // device dictionary
  DevBuf tags, slot_idx, key_ref, id_arena, ctl;
  uint64_t slots = 0, max_keys = 0, arena_cap = 0;
  uint64_t keys_on_host = 0;                // ids already appended to the engine's key table
  void* h_ctl = nullptr;                    // page-locked landing area
};

namespace {
int32_t dfail(sgr_dingest* g, int32_t code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (g) g->last_error = buf;
  return code;
}
#define DG_TRY(g, call)                                                                                          \
  do {                                                                                                           \
    cudaError_t _e = (call);                                                                                     \
    if (_e != cudaSuccess) return dfail((g), _e == cudaErrorMemoryAllocation ? SGR_ERR_OOM : SGR_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
  } while (0)

const char* dg_err_text(uint32_t e) {
  switch (e) {
    case DG_CRC: return "CRC-32C mismatch";
    case DG_LZ4_HEADER: return "bad LZ4 frame header";
    case DG_LZ4_BLOCK: return "LZ4 frame / block truncated or inconsistent";
    case DG_LZ4_SEQUENCE: return "malformed LZ4 sequence";
    case DG_LZ4_CHECKSUM: return "LZ4 checksum mismatch";
    case DG_LZ4_TOO_LARGE: return "LZ4 block decodes past its maximum size";
    case DG_RECORD_LENGTH: return "record length runs past the batch";
    case DG_RECORD_MALFORMED: return "record is malformed";
    case DG_RECORD_COUNT: return "recordsCount does not fit the batch";
    case DG_VALUE_LENGTH: return "packed event value outside 8..56 bytes (u32 type, u32 seq, payload)";
    case DG_ID_LENGTH: return "aggregate id too long";
    case DG_STRAY_BYTES: return "stray bytes after the last record";
  }
  return "unknown";
}

void discard_poll(sgr_dingest* g) {
  g->wire.used = 0; g->batches.clear(); g->n_record_slots = 0; g->staged = g->parts; g->poll = sgr_ingest_stats{};
}
}  // namespace

extern "C" {

int32_t sgr_dingest_create(sgr_engine* e, uint64_t max_keys, uint64_t max_id_bytes, sgr_dingest** out) {
  if (!e || !out || !max_keys) return SGR_ERR_INVALID;
  *out = nullptr;
  void* st = nullptr;
  if (sgr_stream(e, &st) != SGR_OK) return SGR_ERR_INVALID;
  sgr_dingest* g = new sgr_dingest();
  g->eng = e; g->stream = (cudaStream_t)st;
  g->max_keys = max_keys;
  g->slots = 1024;
  while (g->slots < 2 * max_keys) g->slots *= 2;          // load factor <= 0.5
  g->arena_cap = max_id_bytes ? max_id_bytes : 32 * max_keys;
  cudaError_t ce;
  if ((ce = g->tags.reserve(g->slots * 8)) != cudaSuccess || (ce = g->slot_idx.reserve(g->slots * 4)) != cudaSuccess ||
      (ce = g->key_ref.reserve(max_keys * 8)) != cudaSuccess || (ce = g->id_arena.reserve(g->arena_cap + 64)) != cudaSuccess ||
      (ce = g->ctl.reserve(64)) != cudaSuccess || (ce = cudaHostAlloc(&g->h_ctl, 256, cudaHostAllocDefault)) != cudaSuccess ||
      (ce = cudaMemsetAsync(g->tags.p, 0, g->slots * 8, g->stream)) != cudaSuccess || (ce = cudaMemsetAsync(g->slot_idx.p, 0, g->slots * 4, g->stream)) != cudaSuccess ||
      (ce = cudaMemsetAsync(g->ctl.p, 0, 64, g->stream)) != cudaSuccess || (ce = cudaStreamSynchronize(g->stream)) != cudaSuccess) {
    sgr_dingest_destroy(g);
    return ce == cudaErrorMemoryAllocation ? SGR_ERR_OOM : SGR_ERR_CUDA;
  }
  *out = g;
  return SGR_OK;
}

int32_t sgr_dingest_destroy(sgr_dingest* g) {
  if (!g) return SGR_OK;
  g->wire.b.release(); g->d_batches.release(); g->arena.release(); g->rec_off.release(); g->rec_batch.release(); g->out.release();
  g->tags.release(); g->slot_idx.release(); g->key_ref.release(); g->id_arena.release(); g->ctl.release();
  if (g->h_ctl) cudaFreeHost(g->h_ctl);
  delete g;
  return SGR_OK;
}

const char* sgr_dingest_last_error(const sgr_dingest* g) { return g ? g->last_error.c_str() : "null device-ingest handle"; }

int32_t sgr_dingest_set_null_value_type(sgr_dingest* g, int32_t event_type) {
  if (!g || event_type >= (int32_t)SGR_MAX_TYPES) return dfail(g, SGR_ERR_INVALID, "event type out of range");
  g->null_value_type = event_type < 0 ? -1 : event_type;
  return SGR_OK;
}

int32_t sgr_dingest_set_aborted(sgr_dingest* g, int32_t partition, const int64_t* producer_ids, const int64_t* first_offsets, uint64_t n) {
  if (!g || (n && (!producer_ids || !first_offsets))) return dfail(g, SGR_ERR_INVALID, "null argument");
  PartState& ps = g->staged[partition];
  for (uint64_t i = 0; i < n; ++i) ps.aborted.emplace_back(first_offsets[i], producer_ids[i]);
  std::sort(ps.aborted.begin(), ps.aborted.end());
  return SGR_OK;
}

// Walk the batch headers of one fetch; data batches that a read_committed consumer would deliver are queued for the device.
int32_t sgr_dingest_submit(sgr_dingest* g, int32_t partition, const void* data, uint64_t nbytes, sgr_ingest_stats* stats) {
  if (!g || (!data && nbytes)) return dfail(g, SGR_ERR_INVALID, "null argument");
  const uint8_t* buf = (const uint8_t*)data;
  PartState ps = g->staged[partition];   // work on a copy: a malformed fetch leaves the staged view untouched
  sgr_ingest_stats st{};
  std::vector<DgBatch> add;
  uint64_t slots = 0, pos = 0;
  while (nbytes - pos >= 12) {
    const int64_t base_offset = (int64_t)be64(buf + pos);
    const int32_t batch_length = (int32_t)be32(buf + pos + 8);
    if (batch_length < (int32_t)(kBatchHeader - 12)) return dfail(g, SGR_ERR_INVALID, "partition %d offset %lld: batch length %d is smaller than a v2 header", partition, (long long)base_offset, batch_length);
    const uint64_t total = 12ull + (uint32_t)batch_length;
    if (nbytes - pos < total) break;   // a trailing partial batch: the next fetch repeats it
    const uint8_t* b = buf + pos;
    if ((int8_t)b[16] != 2) return dfail(g, SGR_ERR_UNSUPPORTED, "partition %d offset %lld: message format v%d (only RecordBatch magic 2 is decoded)", partition, (long long)base_offset, (int)(int8_t)b[16]);
    const uint16_t attrs = be16(b + 21);
    const int32_t last_offset_delta = (int32_t)be32(b + 23);
    const int64_t producer_id = (int64_t)be64(b + 43);
    const int32_t records_count = (int32_t)be32(b + 57);
    if (last_offset_delta < 0 || records_count < 0) return dfail(g, SGR_ERR_INVALID, "partition %d offset %lld: negative lastOffsetDelta / recordsCount", partition, (long long)base_offset);
    const int64_t last_offset = base_offset + last_offset_delta;
    const int codec = attrs & 7;
    const bool transactional = attrs & 0x10, control = attrs & 0x20;
    ++st.n_batches;
    while (!ps.aborted.empty() && ps.aborted.front().first <= last_offset) { ps.aborting.insert(ps.aborted.front().second); ps.aborted.erase(ps.aborted.begin()); }
    if (control) {
      // tiny and never compressed by the broker: read on the host (CRC included), it only steers the bookkeeping
      ++st.n_control_batches;
      if (sgr_crc32c(b + 21, total - 21) != be32(b + 17)) return dfail(g, SGR_ERR_INVALID, "partition %d offset %lld: CRC-32C mismatch in a control batch", partition, (long long)base_offset);
      if (codec == 0 && total >= kBatchHeader + 8) {
        // record: varint length, attributes, varlong ts delta, varint offset delta, varint key length, key = int16 version, int16 type
        const uint8_t* r = b + kBatchHeader; const uint8_t* end = b + total;
        auto skip_varint = [&]() { while (r < end && (*r & 0x80)) ++r; if (r < end) ++r; };
        skip_varint(); if (r < end) ++r; skip_varint(); skip_varint();
        int32_t kl = 0; { uint32_t v = 0; int sh = 0; while (r < end) { const uint8_t c = *r++; v |= (uint32_t)(c & 0x7f) << sh; if (!(c & 0x80)) break; sh += 7; } kl = (int32_t)(v >> 1) ^ -(int32_t)(v & 1); }
        if (kl >= 4 && r + 4 <= end && be16(r + 2) == 0) ps.aborting.erase(producer_id);   // ABORT marker ends the transaction
      }
    } else if (transactional && ps.aborting.count(producer_id)) {
      ++st.n_aborted_batches; st.n_aborted_records += (uint64_t)records_count;
    } else if (!(ps.seen && last_offset < ps.decoded_next)) {   // (a batch entirely below the position is all duplicates)
      if (codec != 0 && codec != 3) return dfail(g, SGR_ERR_UNSUPPORTED, "partition %d offset %lld: compression codec %d (none and lz4 are decoded)", partition, (long long)base_offset, codec);
      if (total < kBatchHeader) return dfail(g, SGR_ERR_INVALID, "partition %d offset %lld: batch shorter than its header", partition, (long long)base_offset);
      DgBatch d{};
      d.src_off = g->wire.used + pos; d.base_offset = base_offset; d.min_offset = ps.seen ? ps.decoded_next : INT64_MIN;
      d.total_len = (uint32_t)total; d.n_records = (uint32_t)records_count; d.codec = (uint32_t)codec; d.stored_crc = be32(b + 17);
      d.rec_base = (uint32_t)(g->n_record_slots + slots);
      slots += (uint64_t)records_count;
      if (codec == 3) st.n_compressed_bytes += total - kBatchHeader;
      add.push_back(d);
    } else {
      st.n_duplicates += (uint64_t)records_count;
    }
    if (!ps.seen || last_offset + 1 > ps.decoded_next) ps.decoded_next = last_offset + 1;
    ps.seen = true;
    pos += total;
  }
  st.n_bytes = pos; st.n_trailing_bytes = nbytes - pos;
  if (g->n_record_slots + slots >= (1ull << 32)) return dfail(g, SGR_ERR_CAPACITY, "more than 2^32 records in one poll");
  if (pos) {
    DG_TRY(g, g->wire.ensure(pos + 16, g->stream));
    DG_TRY(g, cudaMemcpyAsync((uint8_t*)g->wire.b.p + g->wire.used, buf, pos, cudaMemcpyHostToDevice, g->stream));
    g->wire.used += (pos + 15) & ~15ull;
  }
  g->batches.insert(g->batches.end(), add.begin(), add.end());
  g->n_record_slots += slots;
  g->staged[partition] = ps;
  sgr_ingest_stats& t = g->poll;
  t.n_bytes += st.n_bytes; t.n_batches += st.n_batches; t.n_control_batches += st.n_control_batches; t.n_aborted_batches += st.n_aborted_batches;
  t.n_aborted_records += st.n_aborted_records; t.n_duplicates += st.n_duplicates; t.n_compressed_bytes += st.n_compressed_bytes;
  t.n_trailing_bytes = st.n_trailing_bytes;
  if (stats) *stats = st;
  return SGR_OK;
}

int32_t sgr_dingest_fold(sgr_dingest* g, sgr_ingest_stats* stats) {
  if (!g) return SGR_ERR_INVALID;
  const uint32_t nb = (uint32_t)g->batches.size();
  const uint32_t nrec = (uint32_t)g->n_record_slots;
  sgr_ingest_stats st = g->poll;
  unsigned long long* h = (unsigned long long*)g->h_ctl;
  if (nb) {
    DG_TRY(g, g->d_batches.reserve((size_t)nb * sizeof(DgBatch)));
    DG_TRY(g, cudaMemcpyAsync(g->d_batches.p, g->batches.data(), (size_t)nb * sizeof(DgBatch), cudaMemcpyHostToDevice, g->stream));
    DG_TRY(g, dg_launch_crc_size((const uint8_t*)g->wire.b.p, (DgBatch*)g->d_batches.p, nb, g->stream));
    DG_TRY(g, cudaMemcpyAsync(g->batches.data(), g->d_batches.p, (size_t)nb * sizeof(DgBatch), cudaMemcpyDeviceToHost, g->stream));
    DG_TRY(g, cudaStreamSynchronize(g->stream));
    uint64_t arena_need = 0;
    for (uint32_t i = 0; i < nb; ++i) {
      DgBatch& b = g->batches[i];
      if (b.err) { const int32_t rc = dfail(g, SGR_ERR_INVALID, "offset %lld: %s", (long long)b.base_offset, dg_err_text(b.err)); discard_poll(g); return rc; }
      if (b.codec == 3) { b.arena_off = arena_need; arena_need += ((uint64_t)b.dsize + 15) & ~15ull; st.n_decompressed_bytes += b.dsize; }
    }
    DG_TRY(g, g->arena.reserve(arena_need + 64));
    DG_TRY(g, g->rec_off.reserve((size_t)nrec * 4 + 64));
    DG_TRY(g, g->rec_batch.reserve((size_t)nrec * 4 + 64));
    DG_TRY(g, g->out.reserve((size_t)nrec * 64 + 64));
    DG_TRY(g, cudaMemsetAsync(g->rec_batch.p, 0xff, (size_t)nrec * 4 + 4, g->stream));
    DG_TRY(g, cudaMemcpyAsync(g->d_batches.p, g->batches.data(), (size_t)nb * sizeof(DgBatch), cudaMemcpyHostToDevice, g->stream));
    DG_TRY(g, dg_launch_decode_walk((const uint8_t*)g->wire.b.p, (uint8_t*)g->arena.p, (DgBatch*)g->d_batches.p, nb, (uint32_t*)g->rec_off.p, (uint32_t*)g->rec_batch.p, g->stream));
    // per-poll counters: [2] markers [3] null values [4] duplicates [6] records written; [0] keys / [1] arena / [5] overflow persist
    DG_TRY(g, cudaMemcpyAsync(h, g->ctl.p, 64, cudaMemcpyDeviceToHost, g->stream));
    DG_TRY(g, cudaStreamSynchronize(g->stream));
    const unsigned long long keys_before = h[0], arena_before = h[1];
    h[2] = h[3] = h[4] = h[5] = h[6] = 0;
    DG_TRY(g, cudaMemcpyAsync(g->ctl.p, h, 64, cudaMemcpyHostToDevice, g->stream));
    DgParse p{};
    p.wire = (const uint8_t*)g->wire.b.p; p.arena = (const uint8_t*)g->arena.p; p.batches = (DgBatch*)g->d_batches.p; p.n_batches = nb; p.n_records = nrec;
    p.rec_off = (const uint32_t*)g->rec_off.p; p.rec_batch = (const uint32_t*)g->rec_batch.p; p.out = (uint8_t*)g->out.p; p.null_value_type = g->null_value_type;
    p.dict.tags = (unsigned long long*)g->tags.p; p.dict.slot_idx = (uint32_t*)g->slot_idx.p; p.dict.key_ref = (uint2*)g->key_ref.p;
    p.dict.arena = (uint8_t*)g->id_arena.p; p.dict.ctl = (unsigned long long*)g->ctl.p; p.dict.slots_mask = g->slots - 1;
    p.dict.max_keys = g->max_keys; p.dict.arena_cap = g->arena_cap;
    DG_TRY(g, dg_launch_parse(p, g->stream));
    DG_TRY(g, cudaMemcpyAsync(g->batches.data(), g->d_batches.p, (size_t)nb * sizeof(DgBatch), cudaMemcpyDeviceToHost, g->stream));
    DG_TRY(g, cudaMemcpyAsync(h, g->ctl.p, 64, cudaMemcpyDeviceToHost, g->stream));
    DG_TRY(g, cudaStreamSynchronize(g->stream));
    for (uint32_t i = 0; i < nb; ++i)
      if (g->batches[i].err) {
        const int32_t rc = dfail(g, SGR_ERR_INVALID, "offset %lld, record %u: %s", (long long)g->batches[i].base_offset, g->batches[i].err_record, dg_err_text(g->batches[i].err));
        // ids interned by this failed poll stay in the dictionary (harmless: an id is an id); the records are dropped
        discard_poll(g); return rc;
      }
    if (h[5]) {
      const int32_t rc = dfail(g, SGR_ERR_CAPACITY, "device id dictionary full (%llu ids / %llu id bytes allowed): create the device ingest with larger bounds", (unsigned long long)g->max_keys, (unsigned long long)g->arena_cap);
      discard_poll(g); return rc;
    }
    st.n_markers = h[2]; st.n_null_values = h[3]; st.n_duplicates += h[4]; st.n_records = h[6]; st.n_new_keys = h[0] - keys_before;
    (void)arena_before;
    // ---- grow the table for the new ids, hand their names to the engine's key table, fold
    const uint64_t n_keys = h[0];
    void* d_states = nullptr; uint64_t n_agg = 0; uint32_t sb = 0;
    const int32_t have = sgr_states_device(g->eng, &d_states, &n_agg, &sb);
    if (have != SGR_OK || n_keys > n_agg) {
      uint64_t cap = have == SGR_OK ? n_agg : 0;
      if (cap < 1024) cap = 1024;
      while (cap < n_keys) cap *= 2;
      if (cap > g->max_keys && g->max_keys >= n_keys) cap = g->max_keys;
      int32_t rc = sgr_grow_states(g->eng, cap);
      if (rc) { dfail(g, rc, "engine: %s", sgr_last_error(g->eng)); discard_poll(g); return rc; }
    }
    if (n_keys > g->keys_on_host) {
      const uint64_t add = n_keys - g->keys_on_host;
      std::vector<uint2> refs(add);
      DG_TRY(g, cudaMemcpyAsync(refs.data(), (uint2*)g->key_ref.p + g->keys_on_host, add * 8, cudaMemcpyDeviceToHost, g->stream));
      std::vector<uint8_t> ar(h[1]);
      DG_TRY(g, cudaMemcpyAsync(ar.data(), g->id_arena.p, h[1], cudaMemcpyDeviceToHost, g->stream));   // (whole arena: new ids are scattered in it)
      DG_TRY(g, cudaStreamSynchronize(g->stream));
      std::vector<uint8_t> bytes; std::vector<uint32_t> offs(add + 1, 0);
      for (uint64_t i = 0; i < add; ++i) {
        bytes.insert(bytes.end(), ar.data() + ((uint64_t)refs[i].x << 3), ar.data() + ((uint64_t)refs[i].x << 3) + refs[i].y);
        offs[i + 1] = (uint32_t)bytes.size();
      }
      int32_t rc = sgr_append_keys(g->eng, g, bytes.data(), offs.data(), add);
      if (rc) { dfail(g, rc, "engine: %s", sgr_last_error(g->eng)); discard_poll(g); return rc; }
      g->keys_on_host = n_keys;
    }
    if (nrec) {
      int32_t rc = sgr_fold_incremental_device(g->eng, g->out.p, nrec);
      if (rc) { dfail(g, rc, "engine: %s", sgr_last_error(g->eng)); discard_poll(g); return rc; }
    }
  }
  // ---- commit: the staged positions become the live ones and everything decoded is folded
  for (auto& kv : g->staged) { kv.second.folded_next = kv.second.decoded_next; }
  g->parts = g->staged;
  g->wire.used = 0; g->batches.clear(); g->n_record_slots = 0; g->poll = sgr_ingest_stats{};
  sgr_ingest_stats& t = g->total;
  t.n_bytes += st.n_bytes; t.n_batches += st.n_batches; t.n_records += st.n_records; t.n_markers += st.n_markers; t.n_null_values += st.n_null_values;
  t.n_control_batches += st.n_control_batches; t.n_aborted_batches += st.n_aborted_batches; t.n_aborted_records += st.n_aborted_records;
  t.n_duplicates += st.n_duplicates; t.n_new_keys += st.n_new_keys; t.n_compressed_bytes += st.n_compressed_bytes; t.n_decompressed_bytes += st.n_decompressed_bytes;
  if (stats) *stats = st;
  return SGR_OK;
}

int32_t sgr_dingest_offsets(sgr_dingest* g, int32_t partition, int64_t* decoded_next, int64_t* folded_next) {
  if (!g) return SGR_ERR_INVALID;
  auto it = g->parts.find(partition);
  if (decoded_next) *decoded_next = it == g->parts.end() ? 0 : it->second.decoded_next;
  if (folded_next) *folded_next = it == g->parts.end() ? 0 : it->second.folded_next;
  return SGR_OK;
}

int32_t sgr_dingest_get_stats(sgr_dingest* g, sgr_ingest_stats* out) {
  if (!g || !out) return SGR_ERR_INVALID;
  *out = g->total;
  return SGR_OK;
}

}  // extern "C"
