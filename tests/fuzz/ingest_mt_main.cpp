// ThreadSanitizer harness for the multi-threaded decode (tests/test_ingest_fuzz.py builds it with -fsanitize=thread).
// Input file: u32 n_polls, then per poll: u32 n_fetches, then per fetch: i32 partition, u32 length, bytes.
// Every poll is decoded twice — by one thread and by `threads` threads, into two ingests — and both must agree on the
// pending records, the id dictionary and the offsets. Any data race is a TSan report (non-zero exit).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/sgr.h"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const uint32_t threads = (uint32_t)atoi(argv[2]);
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  uint32_t n_polls = 0;
  if (fread(&n_polls, 4, 1, f) != 1) return 2;
  sgr_ingest *a = nullptr, *b = nullptr;
  if (sgr_ingest_create(&a) || sgr_ingest_create(&b)) return 2;
  uint64_t total = 0;
  for (uint32_t p = 0; p < n_polls; ++p) {
    uint32_t n = 0;
    if (fread(&n, 4, 1, f) != 1) return 2;
    std::vector<std::vector<uint8_t>> bufs(n);
    std::vector<int32_t> parts(n);
    std::vector<const void*> ptrs(n);
    std::vector<uint64_t> lens(n);
    for (uint32_t i = 0; i < n; ++i) {
      uint32_t len;
      if (fread(&parts[i], 4, 1, f) != 1 || fread(&len, 4, 1, f) != 1) return 2;
      bufs[i].resize(len);
      if (len && fread(bufs[i].data(), 1, len, f) != len) return 2;
      ptrs[i] = bufs[i].data(); lens[i] = len;
    }
    const int32_t ra = sgr_ingest_record_batches_mt(a, n, parts.data(), ptrs.data(), lens.data(), 1, nullptr);
    const int32_t rb = sgr_ingest_record_batches_mt(b, n, parts.data(), ptrs.data(), lens.data(), threads, nullptr);
    if (ra != rb) { fprintf(stderr, "poll %u: status %d vs %d\n", p, ra, rb); return 1; }
    const void *pa, *pb; uint64_t na, nb;
    sgr_ingest_pending(a, &pa, &na); sgr_ingest_pending(b, &pb, &nb);
    if (na != nb || (na && memcmp(pa, pb, na * 64))) { fprintf(stderr, "poll %u: pending logs differ\n", p); return 1; }
    const uint8_t *ka, *kb; const uint32_t *oa, *ob; uint64_t ca, cb;
    sgr_ingest_keys(a, &ka, &oa, &ca); sgr_ingest_keys(b, &kb, &ob, &cb);
    if (ca != cb || memcmp(oa, ob, (ca + 1) * 4) || (ca && memcmp(ka, kb, oa[ca]))) { fprintf(stderr, "poll %u: dictionaries differ\n", p); return 1; }
    total += na;
    if (p % 2) { sgr_ingest_mark_folded(a); sgr_ingest_mark_folded(b); }
  }
  sgr_ingest_destroy(a); sgr_ingest_destroy(b);
  fclose(f);
  printf("polls=%u records=%llu\n", n_polls, (unsigned long long)total);
  return 0;
}
