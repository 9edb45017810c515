// lz4_fast.h — LZ4 frame decode for ONE thread that never waits for its input (device ingest, dingest_kernels.cu).
//
// Why this shape. A producer batch is one LZ4 block of ~2000 tiny sequences (literals ~2 bytes, matches ~6 bytes on event
// topics): a serial chain, so a batch is a thread and the parallelism is the tens of thousands of batches of a poll. In a warp
// of 32 independent batches every memory access of the chain costs the WHOLE warp a round trip: some lane always misses, and the
// scoreboard that guards a load's destination register is per warp, not per lane — a register "prefetch" by one lane stalls the
// next instruction of any other lane that touches the same register name (measured: profiles/r02b_dingest_fast_v1_ncu.txt, 65 %
// of all stall samples on three window-shift MOVs). So:
//   * input   comes through a policy object. On the device it is a per-thread ring of eight 16-byte chunks in SHARED memory
//             filled by cp.async six chunks ahead (RingIn, dingest_kernels.cu): asynchronous copies have no destination register,
//             and tokens, lengths, offsets and literals are cut out of two shared-memory words. On the host (HostIn) it reads
//             the bytes where they lie.
//   * output  is an 8-byte accumulator (Out8): bytes are merged in with shifts and every append leaves memory up to date (one
//             aligned 8-byte store, fire and forget), so a match ALWAYS reads its source from memory with the same three
//             aligned loads — no special path for near, overlapping or straddling matches that would serialise the warp
//             (offset < 8 replicates the period in registers). One dependent round trip per sequence remains: the match source.
// Accept / reject behaviour mirrors lz4_frame_decode of ingest.cpp decision for decision (tests/test_lz4_fast_cpu.py runs both on
// the same corpus under ASan; tests/test_gpu_dingest.py compares the device result with the host decoder).
//
// Host-compilable: the same source is built by g++ for the CPU tests. Memory contract: `out` is 8-byte aligned, is written up to
// the next 8-byte boundary past the decoded size and may be READ up to 24 bytes past it; the input policy states its own slack.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LZF_HD __host__ __device__ __forceinline__
#define LZF_HDN __host__ __device__
#else
#define LZF_HD inline
#define LZF_HDN inline
#endif

namespace sgr {
namespace lzf {

enum : uint32_t {   // numerically the DgErr codes of dingest_kernels.cuh
  OK = 0, HEADER = 2, BLOCK = 3, SEQUENCE = 4, CHECKSUM = 5, TOO_LARGE = 6,
};

LZF_HD uint64_t low_bytes(uint64_t v, uint32_t k) { return k >= 8 ? v : v & ((1ull << (k * 8)) - 1); }
LZF_HD uint64_t funnel(uint64_t a, uint64_t b, uint32_t byte_shift) {   // bytes [byte_shift, byte_shift + 8) of the 16 bytes a:b
  const uint32_t sh = byte_shift * 8;
  return sh ? (a >> sh) | (b << (64 - sh)) : a;
}
LZF_HD uint32_t rd32(const uint8_t* p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

LZF_HD uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
LZF_HDN uint32_t xxh32(const uint8_t* p, uint64_t len, uint32_t seed) {
  const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
  const uint8_t* end = p + len;
  uint32_t h;
  if (len >= 16) {
    const uint8_t* limit = end - 16;
    uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    do {
      v1 = rotl(v1 + rd32(p) * P2, 13) * P1; p += 4;
      v2 = rotl(v2 + rd32(p) * P2, 13) * P1; p += 4;
      v3 = rotl(v3 + rd32(p) * P2, 13) * P1; p += 4;
      v4 = rotl(v4 + rd32(p) * P2, 13) * P1; p += 4;
    } while (p <= limit);
    h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
  } else {
    h = seed + P5;
  }
  h += (uint32_t)len;
  while (p + 4 <= end) { h = rotl(h + rd32(p) * P3, 17) * P4; p += 4; }
  while (p < end) { h = rotl(h + (*p++) * P5, 11) * P1; }
  h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
  return h;
}

LZF_HD uint64_t ld64(const uint8_t* p) {   // p is 8-byte aligned
#if defined(__CUDA_ARCH__)
  return *reinterpret_cast<const unsigned long long*>(p);
#else
  uint64_t v; memcpy(&v, p, 8); return v;
#endif
}
LZF_HD void st64(uint8_t* p, uint64_t v) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<unsigned long long*>(p) = v;
#else
  memcpy(p, &v, 8);
#endif
}

// Input policy of the host build (and the reference for what a policy provides): seek(p) before the first read of a stream,
// advance(p) before get64(p) whenever p moved forward, get64(p) = the 8 bytes at p (may read 7 bytes past the last one used).
struct HostIn {
  LZF_HD void seek(const uint8_t*) {}
  LZF_HD void advance(const uint8_t*) {}
  LZF_HD uint64_t get64(const uint8_t* p) const { uint64_t v; memcpy(&v, p, 8); return v; }
};

// The output: the 8-byte word under construction; memory holds every byte below op after every call.
struct Out8 {
  uint8_t* out;
  uint64_t op;
  uint64_t acc;
  bool src_bypass_l1;   // device: read match sources with ld.global.cg
  LZF_HD void init(uint8_t* o) { out = o; op = 0; acc = 0; src_bypass_l1 = false; }
  LZF_HD uint64_t src64(const uint8_t* p) const {
#if defined(__CUDA_ARCH__)
    if (src_bypass_l1) return __ldcg(reinterpret_cast<const unsigned long long*>(p));
#endif
    return ld64(p);
  }
  LZF_HD void put(uint64_t v, uint32_t k) {   // append k (1..8) bytes; v is zero above them
    const uint32_t q = (uint32_t)op & 7, sh = q * 8;
    uint8_t* word = out + (op & ~7ull);
    acc |= v << sh;
    op += k;
    if (q + k >= 8) { st64(word, acc); acc = sh ? v >> (64 - sh) : 0; word += 8; }
    if (op & 7) st64(word, acc);
  }
  LZF_HD static uint64_t replicate(uint64_t p, uint32_t period) {   // p: `period` (1..7) bytes -> 8 bytes of their repetition
    for (uint32_t n = period; n < 8; n <<= 1) p |= p << (8 * n);
    return p;
  }
  LZF_HD void match(uint32_t off, uint64_t len) {   // append `len` bytes that repeat the output `off` bytes back
    while (len) {
      const uint64_t s = op - off;
      const uint8_t* a = out + (s & ~7ull);
      const uint32_t i = (uint32_t)s & 7;
      const uint64_t w0 = src64(a), w1 = src64(a + 8);
      const uint64_t w2 = (off >= 16 && len > 8) ? src64(a + 16) : 0;   // (a load nobody needs still costs the warp its wavefronts)
      uint64_t v = funnel(w0, w1, i);
      if (off < 8) v = replicate(low_bytes(v, off), off);
      uint32_t k = len < 8 ? (uint32_t)len : 8u;
      put(low_bytes(v, k), k);
      len -= k;
      if (off >= 16 && len) {   // the second 8 source bytes were in memory before this step too
        k = len < 8 ? (uint32_t)len : 8u;
        put(low_bytes(funnel(w1, w2, i), k), k);
        len -= k;
      }
    }
  }
  template <class IN>
  LZF_HD void literals(IN& in, const uint8_t* p, uint64_t len) {
    while (len) {
      in.advance(p);
      const uint32_t k = len < 8 ? (uint32_t)len : 8u;
      put(low_bytes(in.get64(p), k), k);
      p += k; len -= k;
    }
  }
};

// DECODE = false: validate and measure only (no output is touched). Returns an error code (enum above); *out_len = decoded bytes.
template <bool DECODE, class IN>
LZF_HDN uint32_t frame(IN& in, const uint8_t* src, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_len, bool src_bypass_l1 = false) {
  if (n < 7) return HEADER;
  if (rd32(src) != 0x184D2204u) return HEADER;
  const uint8_t flg = src[4], bd = src[5];
  if ((flg >> 6) != 1 || (flg & 0x02)) return HEADER;
  const bool block_checksum = flg & 0x10, content_size = flg & 0x08, content_checksum = flg & 0x04, dict_id = flg & 0x01;
  const uint32_t bs_code = (bd >> 4) & 7;
  if (bs_code < 4 || (bd & 0x8F)) return HEADER;
  const uint64_t max_block = 1ull << (8 + 2 * bs_code);
  const uint64_t desc_len = 2 + (content_size ? 8 : 0) + (dict_id ? 4 : 0);
  if (n < 4 + desc_len + 1) return HEADER;
  uint64_t declared = 0;
  if (content_size) for (int k = 7; k >= 0; --k) declared = (declared << 8) | src[6 + k];
  if (((xxh32(src + 4, desc_len, 0) >> 8) & 0xff) != src[4 + desc_len]) return HEADER;
  uint64_t pos = 4 + desc_len + 1;
  Out8 w;
  w.init(out);
  w.src_bypass_l1 = src_bypass_l1;
  uint64_t op = 0;   // decoded bytes so far (== w.op when DECODE)
  for (;;) {
    if (pos + 4 > n) return BLOCK;
    const uint32_t word = rd32(src + pos); pos += 4;
    if (word == 0) break;
    const bool stored = word & 0x80000000u;
    const uint64_t bsz = word & 0x7FFFFFFFu;
    if (bsz > max_block) return BLOCK;
    if (pos + bsz + (block_checksum ? 4 : 0) > n) return BLOCK;
    const uint8_t* b = src + pos;
    if (block_checksum && xxh32(b, bsz, 0) != rd32(b + bsz)) return CHECKSUM;
    if (stored) {
      if (DECODE) {
        if (op + bsz > out_cap) return TOO_LARGE;
        if (bsz) { in.seek(b); w.literals(in, b, bsz); }
      }
      op += bsz;
    } else {
      const uint64_t block_start = op;
      uint64_t ip = 0;
      in.seek(b);
      for (;;) {
        if (ip >= bsz) return SEQUENCE;
        in.advance(b + ip);
        uint64_t v = in.get64(b + ip);
        const uint32_t token = (uint32_t)v & 0xffu;
        ++ip;
        uint64_t lit = token >> 4;
        bool lits_in_v = lit < 8;   // the literals of a short run sit in the same 8 bytes as the token
        if (lit == 15) {
          uint32_t s;
          do {
            if (ip >= bsz) return SEQUENCE;
            in.advance(b + ip);
            s = (uint32_t)in.get64(b + ip) & 0xffu; ++ip; lit += s;
          } while (s == 255);
        }
        if (lit > bsz - ip) return SEQUENCE;
        if (op - block_start + lit > max_block) return TOO_LARGE;
        if (DECODE && lit) {
          if (op + lit > out_cap) return TOO_LARGE;
          if (lits_in_v) w.put(low_bytes(v >> 8, (uint32_t)lit), (uint32_t)lit);
          else w.literals(in, b + ip, lit);
        }
        op += lit; ip += lit;
        if (ip == bsz) break;   // the last sequence carries literals only
        if (ip + 2 > bsz) return SEQUENCE;
        in.advance(b + ip);
        v = in.get64(b + ip);
        const uint32_t off = (uint32_t)v & 0xffffu; ip += 2;
        uint64_t mlen = token & 15;
        if (mlen == 15) {
          uint32_t s;
          do {
            if (ip >= bsz) return SEQUENCE;
            in.advance(b + ip);
            s = (uint32_t)in.get64(b + ip) & 0xffu; ++ip; mlen += s;
          } while (s == 255);
        }
        mlen += 4;
        if (off == 0 || off > op) return SEQUENCE;   // matches may reach back across blocks, never before the frame
        if (op - block_start + mlen > max_block) return TOO_LARGE;
        if (DECODE) {
          if (op + mlen > out_cap) return TOO_LARGE;
          w.match(off, mlen);
        }
        op += mlen;
      }
    }
    pos += bsz + (block_checksum ? 4 : 0);
  }
  if (content_checksum) {
    if (pos + 4 > n) return BLOCK;
    if (DECODE && xxh32(out, op, 0) != rd32(src + pos)) return CHECKSUM;
    pos += 4;
  }
  if (content_size && declared != op) return BLOCK;
  *out_len = op;
  return OK;
}

}  // namespace lzf
}  // namespace sgr
