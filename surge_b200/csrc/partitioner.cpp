// partitioner.cpp — aggregate id -> Kafka partition, host side.
//
// KafkaPartitionProvider.partitionForKey(s, n) = math.abs(MurmurHash3.stringHash(s) % n)
//   modules/common/src/main/scala/surge/kafka/KafkaPartitioner.scala:7-9
// over PartitionStringUpToColon: s.takeWhile(_ != ':')           (same file, :38-42)
// The hash is scala-library 2.13.8's MurmurHash3.stringHash over UTF-16 code units (a
// third-party dependency of the reference; the reference holds no known-answer vector for
// it, so ownership parity is "unpinned" against Scala — tests/test_partition_hash_pin.py pins this code to a real
// MurmurHash3_x86_32 instead). Folded state bytes never depend
// on it; it only decides which rank owns an aggregate.
#include <stdint.h>

#include <vector>

#include "../../include/sgr.h"

namespace {
inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
inline uint32_t mix_last(uint32_t h, uint32_t k) {
  k *= 0xcc9e2d51u; k = rotl(k, 15); k *= 0x1b873593u;
  return h ^ k;
}
inline uint32_t mix(uint32_t h, uint32_t k) {
  h = mix_last(h, k); h = rotl(h, 13);
  return h * 5u + 0xe6546b64u;
}
inline uint32_t finalize(uint32_t h, uint32_t len) {
  h ^= len;
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
}  // namespace

extern "C" int32_t sgr_string_hash_utf16(const uint16_t* u, uint32_t n) {
  uint32_t h = 0xf7ca7fd2u;  // MurmurHash3.stringSeed
  uint32_t i = 0;
  for (; i + 1 < n; i += 2) h = mix(h, ((uint32_t)u[i] << 16) + (uint32_t)u[i + 1]);
  if (i < n) h = mix_last(h, (uint32_t)u[i]);
  return (int32_t)finalize(h, n);
}

extern "C" int32_t sgr_partition_for_key_utf8(const uint8_t* key, uint32_t klen, uint32_t num_partitions,
                                              int32_t up_to_colon, int32_t* partition) {
  if ((!key && klen) || !partition || num_partitions == 0 || num_partitions > 0x7fffffffu) return SGR_ERR_INVALID;
  // UTF-8 -> UTF-16 code units (what a JVM String holds)
  std::vector<uint16_t> u;
  u.reserve(klen);
  for (uint32_t i = 0; i < klen;) {
    uint32_t c = key[i], cp, extra;
    if (c < 0x80) { cp = c; extra = 0; }
    else if ((c >> 5) == 6) { cp = c & 31u; extra = 1; }
    else if ((c >> 4) == 14) { cp = c & 15u; extra = 2; }
    else if ((c >> 3) == 30) { cp = c & 7u; extra = 3; }
    else return SGR_ERR_INVALID;
    if (i + extra >= klen + 0u && extra) { if (i + extra > klen - 1) return SGR_ERR_INVALID; }
    for (uint32_t j = 1; j <= extra; ++j) {
      if ((key[i + j] >> 6) != 2) return SGR_ERR_INVALID;
      cp = (cp << 6) | (key[i + j] & 63u);
    }
    i += extra + 1;
    if (cp >= 0x10000) { cp -= 0x10000; u.push_back((uint16_t)(0xd800 + (cp >> 10))); u.push_back((uint16_t)(0xdc00 + (cp & 1023))); }
    else u.push_back((uint16_t)cp);
  }
  uint32_t n = (uint32_t)u.size();
  if (up_to_colon) { uint32_t k = 0; while (k < n && u[k] != (uint16_t)':') ++k; n = k; }
  const int32_t h = sgr_string_hash_utf16(u.data(), n);
  int32_t r = h % (int32_t)num_partitions;  // truncating remainder, sign of the dividend, like the JVM
  *partition = r < 0 ? -r : r;
  return SGR_OK;
}

// Vectorised form for a whole key table: partition_of[i] for key i = keys[key_offsets[i] .. key_offsets[i+1]).
extern "C" int32_t sgr_partitions_for_keys(const uint8_t* keys, const uint32_t* key_offsets, uint64_t n, uint32_t num_partitions,
                                           int32_t up_to_colon, uint32_t* partition_of) {
  if ((!keys && n && key_offsets[n]) || !key_offsets || !partition_of) return SGR_ERR_INVALID;
  for (uint64_t i = 0; i < n; ++i) {
    int32_t p = 0;
    const int32_t rc = sgr_partition_for_key_utf8(keys + key_offsets[i], key_offsets[i + 1] - key_offsets[i], num_partitions, up_to_colon, &p);
    if (rc != SGR_OK) return rc;
    partition_of[i] = (uint32_t)p;
  }
  return SGR_OK;
}
