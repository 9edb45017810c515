// Sanitizer harness for the record-batch decoder (tests/test_ingest_fuzz.py builds it with
// g++ -fsanitize=address,undefined together with surge_b200/csrc/ingest.cpp and feeds it a corpus file).
// Corpus: u32 count, then per case: u8 kind (0 = record batches, 1 = lz4 frame, 2 = batches of JSON values,
// 3 = batches of protobuf-wrapped values), u32 length, bytes.
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
  // kind 2: well-framed batches whose VALUES are (mutated) JSON, decoded by an ingest with a JSON packer;
  // kind 3: the same through the protobuf Event unwrapping
  sgr_ingest *js = nullptr, *pb = nullptr;
  if (sgr_ingest_create(&js) != SGR_OK || sgr_ingest_create(&pb) != SGR_OK) return 2;
  sgr_json_event ev[2];
  memset(ev, 0, sizeof ev);
  ev[0].type_name = "Inc"; ev[0].event_type = 0; ev[0].n_fields = 3;
  ev[0].fields[0].name = "by"; ev[0].fields[0].kind = SGR_JSON_I32; ev[0].fields[0].dst_off = 16;
  ev[0].fields[1].name = "seq"; ev[0].fields[1].kind = SGR_JSON_I32; ev[0].fields[1].dst_off = 4;
  ev[0].fields[2].name = "w"; ev[0].fields[2].kind = SGR_JSON_F64; ev[0].fields[2].dst_off = 56;
  ev[1].type_name = "Big\xc3\xa9"; ev[1].event_type = 1; ev[1].n_fields = 1;
  ev[1].fields[0].name = "v"; ev[1].fields[0].kind = SGR_JSON_I64; ev[1].fields[0].dst_off = 24;
  if (sgr_ingest_set_json_packer(js, "_type", ev, 2, 5) != SGR_OK || sgr_ingest_set_value_framing(js, SGR_VALUE_JSON) != SGR_OK) return 2;
  if (sgr_ingest_set_value_framing(pb, SGR_VALUE_PROTOBUF_EVENT) != SGR_OK) return 2;
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
    } else if (kind == 2 || kind == 3) {
      sgr_ingest* g = kind == 2 ? js : pb;
      sgr_ingest_stats st;
      rc = sgr_ingest_record_batches(g, (int32_t)i, buf, len, &st);              // a fresh partition per case: nothing is a duplicate
      sgr_ingest_mark_folded(g);
    } else {
      sgr_ingest_stats st;
      rc = sgr_ingest_record_batches(shared, (int32_t)(i % 3), buf, len, &st);   // state carries over between cases
      if (i % 64 == 63) sgr_ingest_mark_folded(shared);
    }
    (rc == SGR_OK ? ok : refused) += 1;
    delete[] buf;
  }
  sgr_ingest_destroy(shared); sgr_ingest_destroy(js); sgr_ingest_destroy(pb);
  fclose(f);
  printf("cases=%u ok=%llu refused=%llu\n", count, (unsigned long long)ok, (unsigned long long)refused);
  return 0;
}
