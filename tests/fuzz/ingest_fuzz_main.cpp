// Sanitizer harness for the record-batch decoder (tests/test_ingest_fuzz.py builds it with
// g++ -fsanitize=address,undefined together with surge_b200/csrc/ingest.cpp and feeds it a corpus file).
// Corpus: u32 count, then per case: u8 kind (0 = record batches, 1 = lz4 frame), u32 length, bytes.
// Every case must come back with a status code — never a crash, an out-of-bounds access or undefined behaviour.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/sgr.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  uint32_t count = 0;
  if (fread(&count, 4, 1, f) != 1) return 2;
  sgr_ingest* shared = nullptr;
  if (sgr_ingest_create(&shared) != SGR_OK) return 2;
  uint64_t ok = 0, refused = 0;
  std::vector<uint8_t> out(1 << 20);
  for (uint32_t i = 0; i < count; ++i) {
    uint8_t kind; uint32_t len;
    if (fread(&kind, 1, 1, f) != 1 || fread(&len, 4, 1, f) != 1) return 2;
    // exact-size heap buffer: any read past the end is an ASan report
    uint8_t* buf = new uint8_t[len ? len : 1];
    if (len && fread(buf, 1, len, f) != len) return 2;
    int32_t rc;
    if (kind == 1) {
      uint64_t n = 0;
      rc = sgr_lz4_frame_decode(buf, len, out.data(), out.size(), &n);
    } else {
      sgr_ingest_stats st;
      rc = sgr_ingest_record_batches(shared, (int32_t)(i % 3), buf, len, &st);   // state carries over between cases
      if (i % 64 == 63) sgr_ingest_mark_folded(shared);
    }
    (rc == SGR_OK ? ok : refused) += 1;
    delete[] buf;
  }
  sgr_ingest_destroy(shared);
  fclose(f);
  printf("cases=%u ok=%llu refused=%llu\n", count, (unsigned long long)ok, (unsigned long long)refused);
  return 0;
}
