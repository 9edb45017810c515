/* kafka_encode.c — TEST INFRASTRUCTURE: a fast producer-side encoder of Kafka RecordBatch v2 streams for the Counter events topic.
 *
 * bench.py's end-to-end leg and tests/test_gpu_dingest.py need tens of millions of records in the byte format a broker hands to a
 * consumer; oracle/kafka_batch.py (pure Python, the readable restatement) manages thousands per second. This file writes the
 * same bytes at memory speed. It is pinned, not trusted: tests/test_kafka_encode.py decodes its output with oracle/kafka_batch.py,
 * with the native host decoder (csrc/ingest.cpp) and its lz4 frames with liblz4 (pyarrow), and compares record for record.
 *
 * Format restated (org.apache.kafka:kafka-clients:3.2.3, third-party — byte-level parity with a real broker is UNPINNED, see
 * oracle/kafka_batch.py's header): DefaultRecordBatch header (61 bytes, big endian, CRC-32C over attributes..end), records as
 * zig-zag varints (DefaultRecord), compression none or lz4 (frame: magic 0x184D2204, FLG 0x60 = version 01 + independent blocks,
 * BD 0x40 = 64 KiB blocks, header checksum = second byte of xxHash32 of the descriptor; KafkaLZ4BlockOutputStream).
 * Record key  = "agg-<n>:<seq>"  (s"${evt.aggregateId}:${evt.sequenceNumber}", core TestBoundedContext.scala:159-161)
 * Record value = u32 type, u32 seq, i32 by (little endian): the packed Counter event of surge_b200/formats.py.
 * Only tests/, bench.py's input construction and __graft_entry__.smoke() may call this. */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint32_t crc_tab[8][256];
static int crc_ready = 0;
static void crc_init(void) {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
    crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; i++)
    for (int k = 1; k < 8; k++) crc_tab[k][i] = (crc_tab[k - 1][i] >> 8) ^ crc_tab[0][crc_tab[k - 1][i] & 0xff];
  crc_ready = 1;
}
static uint32_t crc32c(const uint8_t* p, uint64_t n) {
  if (!crc_ready) crc_init();
  uint32_t crc = 0xffffffffu;
  while (n >= 8) {
    uint64_t w; memcpy(&w, p, 8);
    uint32_t lo = (uint32_t)w ^ crc, hi = (uint32_t)(w >> 32);
    crc = crc_tab[7][lo & 0xff] ^ crc_tab[6][(lo >> 8) & 0xff] ^ crc_tab[5][(lo >> 16) & 0xff] ^ crc_tab[4][lo >> 24] ^
          crc_tab[3][hi & 0xff] ^ crc_tab[2][(hi >> 8) & 0xff] ^ crc_tab[1][(hi >> 16) & 0xff] ^ crc_tab[0][hi >> 24];
    p += 8; n -= 8;
  }
  while (n--) crc = (crc >> 8) ^ crc_tab[0][(crc ^ *p++) & 0xff];
  return ~crc;
}

static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint32_t xxh32(const uint8_t* p, uint64_t len, uint32_t seed) {
  const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
  const uint8_t* end = p + len;
  uint32_t h;
  if (len >= 16) {
    const uint8_t* limit = end - 16;
    uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    do {
      v1 = rotl32(v1 + rd32(p) * P2, 13) * P1; p += 4;
      v2 = rotl32(v2 + rd32(p) * P2, 13) * P1; p += 4;
      v3 = rotl32(v3 + rd32(p) * P2, 13) * P1; p += 4;
      v4 = rotl32(v4 + rd32(p) * P2, 13) * P1; p += 4;
    } while (p <= limit);
    h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
  } else h = seed + P5;
  h += (uint32_t)len;
  while (p + 4 <= end) { h = rotl32(h + rd32(p) * P3, 17) * P4; p += 4; }
  while (p < end) { h = rotl32(h + (*p++) * P5, 11) * P1; }
  h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
  return h;
}

/* LZ4 block compression, greedy with a 4-byte hash (the LZ4 block format's end rules: the last 5 bytes are literals, the last
 * match starts at least 12 bytes before the end). Returns the compressed size; dst must hold n + n/255 + 16 bytes. */
static uint32_t lz4_block(const uint8_t* src, uint32_t n, uint8_t* dst) {
  enum { HLOG = 13 };
  static __thread int32_t table[1 << HLOG];
  for (int i = 0; i < (1 << HLOG); i++) table[i] = -1;
  uint32_t ip = 0, anchor = 0, op = 0;
  const uint32_t mflimit = n > 12 ? n - 12 : 0;
  while (n > 12 && ip < mflimit) {
    const uint32_t seq = rd32(src + ip);
    const uint32_t h = (seq * 2654435761u) >> (32 - HLOG);
    const int32_t ref = table[h];
    table[h] = (int32_t)ip;
    if (ref >= 0 && ip - (uint32_t)ref <= 65535 && rd32(src + ref) == seq) {
      uint32_t mlen = 4;
      const uint32_t limit = n - 5;
      while (ip + mlen < limit && src[ref + mlen] == src[ip + mlen]) mlen++;
      const uint32_t lit = ip - anchor;
      uint8_t* token = dst + op++;
      if (lit >= 15) { *token = 0xF0; uint32_t r = lit - 15; while (r >= 255) { dst[op++] = 255; r -= 255; } dst[op++] = (uint8_t)r; }
      else *token = (uint8_t)(lit << 4);
      memcpy(dst + op, src + anchor, lit); op += lit;
      const uint32_t off = ip - (uint32_t)ref;
      dst[op++] = (uint8_t)off; dst[op++] = (uint8_t)(off >> 8);
      uint32_t ml = mlen - 4;
      if (ml >= 15) { *token |= 15; ml -= 15; while (ml >= 255) { dst[op++] = 255; ml -= 255; } dst[op++] = (uint8_t)ml; }
      else *token |= (uint8_t)ml;
      ip += mlen; anchor = ip;
    } else ip++;
  }
  const uint32_t lit = n - anchor;
  uint8_t* token = dst + op++;
  if (lit >= 15) { *token = 0xF0; uint32_t r = lit - 15; while (r >= 255) { dst[op++] = 255; r -= 255; } dst[op++] = (uint8_t)r; }
  else *token = (uint8_t)(lit << 4);
  memcpy(dst + op, src + anchor, lit); op += lit;
  return op;
}

static uint64_t lz4_frame(const uint8_t* src, uint64_t n, uint8_t* dst) {
  uint64_t op = 0;
  const uint8_t hdr[6] = {0x04, 0x22, 0x4D, 0x18, 0x60, 0x40};
  memcpy(dst, hdr, 6); op = 6;
  dst[op++] = (uint8_t)((xxh32(hdr + 4, 2, 0) >> 8) & 0xff);
  for (uint64_t s = 0; s < n; s += 65536) {
    const uint32_t raw = (uint32_t)(n - s < 65536 ? n - s : 65536);
    const uint32_t c = lz4_block(src + s, raw, dst + op + 4);
    uint32_t word;
    if (c >= raw) { word = raw | 0x80000000u; memcpy(dst + op + 4, src + s, raw); }
    else word = c;
    memcpy(dst + op, &word, 4);
    op += 4 + (word & 0x7fffffffu);
  }
  memset(dst + op, 0, 4); op += 4;
  return op;
}

static uint32_t put_uvar(uint8_t* p, uint64_t z) { uint32_t k = 0; while (z & ~0x7Full) { p[k++] = (uint8_t)((z & 0x7f) | 0x80); z >>= 7; } p[k++] = (uint8_t)z; return k; }
static uint32_t put_varint(uint8_t* p, int32_t v) { return put_uvar(p, (uint32_t)((v << 1) ^ (v >> 31))); }
static uint32_t put_varlong(uint8_t* p, int64_t v) { return put_uvar(p, (uint64_t)((v << 1) ^ (v >> 63))); }
static void be16(uint8_t* p, uint16_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }
static void be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
static void be64(uint8_t* p, uint64_t v) { be32(p, (uint32_t)(v >> 32)); be32(p + 4, (uint32_t)v); }

/* Upper bound of the bytes orc_kafka_encode_counter writes for n records in batches of recs_per_batch. */
uint64_t orc_kafka_encode_bound(uint64_t n, uint32_t recs_per_batch) {
  const uint64_t batches = (n + recs_per_batch - 1) / recs_per_batch + 1;
  return n * 64 + batches * (61 + 64 + (uint64_t)recs_per_batch * 64 / 200 + 32);
}

/* Encodes records i = 0..n-1 (aggregate agg[i], event type[i], seq[i], by[i]) of ONE partition as consecutive RecordBatches
 * of recs_per_batch records starting at base_offset. Returns the bytes written, or -1 when `cap` is too small. */
int64_t orc_kafka_encode_counter(const uint32_t* agg, const uint32_t* type, const uint32_t* seq, const int32_t* by, uint64_t n,
                                 uint32_t recs_per_batch, int lz4, int64_t base_offset, uint8_t* out, uint64_t cap) {
  if (!recs_per_batch) return -1;
  const uint64_t body_cap = (uint64_t)recs_per_batch * 80 + 64;
  uint8_t* body = (uint8_t*)malloc(body_cap);
  uint8_t* comp = (uint8_t*)malloc(body_cap + body_cap / 200 + 64);
  if (!body || !comp) { free(body); free(comp); return -1; }
  uint64_t op = 0;
  const int64_t ts0 = 1600000000000ll;
  for (uint64_t s = 0; s < n; s += recs_per_batch) {
    const uint32_t cnt = (uint32_t)(n - s < recs_per_batch ? n - s : recs_per_batch);
    uint64_t bl = 0;
    for (uint32_t d = 0; d < cnt; d++) {
      const uint64_t i = s + d;
      uint8_t rec[96]; uint32_t r = 0;
      rec[r++] = 0;                                   /* attributes */
      r += put_varlong(rec + r, (int64_t)d);          /* timestampDelta */
      r += put_varint(rec + r, (int32_t)d);           /* offsetDelta */
      char key[40];
      const int kl = snprintf(key, sizeof key, "agg-%u:%u", agg[i], seq[i]);
      r += put_varint(rec + r, kl); memcpy(rec + r, key, (size_t)kl); r += (uint32_t)kl;
      r += put_varint(rec + r, 12);
      memcpy(rec + r, &type[i], 4); memcpy(rec + r + 4, &seq[i], 4); memcpy(rec + r + 8, &by[i], 4); r += 12;
      r += put_varint(rec + r, 0);                    /* headers */
      bl += put_varint(body + bl, (int32_t)r);
      memcpy(body + bl, rec, r); bl += r;
    }
    const uint8_t* payload = body; uint64_t pl = bl;
    if (lz4) { pl = lz4_frame(body, bl, comp); payload = comp; }
    const uint64_t total = 61 + pl;
    if (op + total > cap) { free(body); free(comp); return -1; }
    uint8_t* b = out + op;
    be64(b, (uint64_t)(base_offset + (int64_t)s));
    be32(b + 8, (uint32_t)(total - 12));
    be32(b + 12, 0);                                  /* partitionLeaderEpoch */
    b[16] = 2;                                        /* magic */
    be16(b + 21, (uint16_t)(lz4 ? 3 : 0));            /* attributes */
    be32(b + 23, cnt - 1);                            /* lastOffsetDelta */
    be64(b + 27, (uint64_t)ts0); be64(b + 35, (uint64_t)(ts0 + cnt - 1));
    be64(b + 43, (uint64_t)-1ll); be16(b + 51, (uint16_t)-1); be32(b + 53, (uint32_t)-1);   /* producerId, epoch, baseSequence */
    be32(b + 57, cnt);
    memcpy(b + 61, payload, pl);
    be32(b + 17, crc32c(b + 21, total - 21));
    op += total;
  }
  free(body); free(comp);
  return (int64_t)op;
}
