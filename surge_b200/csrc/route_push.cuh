// route_push.cuh — pipelined route + exchange + fold over peer memory (route_push.cu), used by engine.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "bulk_fold.cuh"
#include "dist.cuh"

namespace sgr {

// measurement knob (sgr_set_option "push_tile"): records per CTA of the push kernel, 256 / 512 / 1024
// "push_pull": 1 = the sender partitions into its OWN buffer and the owner's fold reads the regions over NVLink (remote loads);
//              0 = the sender writes into the owner's buffer (remote stores). The same on every rank.
// "push_staged": 1 = partition through shared memory (contiguous per-owner runs), 0 = position pass + direct copy, -1 = staged iff
//                the regions are written over NVLink (pull == 0)
// "push_fold_blocks_per_sm": grid cap of the fold launches beside the partition kernel; 0 = one tile per CTA (the low-priority fold
//                            then fills what the high-priority partition leaves free)
struct PushTuning { int tile = 512; int pull = 1; int fold_blocks_per_sm = 2; int staged = -1; };   // measured best at N = 2, 4 (scripts/push_ab.py)
PushTuning& push_tuning();

struct PushFoldArgs {
  const RowProgram* prog;
  const BulkLayout* lay;
  void* scratch;                 // bulk_scratch_bytes(lay, n_slots), zero outside a fold
  uint8_t* states;               // n_slots x 16, prior states (all zero for a rebuild)
  uint32_t* err_ids;             // n_slots + 1
  unsigned long long* counters;  // 8 x u64
  uint64_t n_slots;
  uint32_t n_chunks;             // the same on every rank
  bool compact;                  // exchange only the record words the program reads
  bool ordered;                  // positions inside the regions follow the log (needed by the exact replay of throwing slots)
  int num_sms;
};
struct PushRegion { const uint8_t* base; uint32_t count; };
struct PushFoldResult {
  uint64_t n_recv = 0, n_err_slots = 0;
  bool any_err_slots = false;        // some rank of the job saw a throwing slot
  float ms_push = 0, ms_total = 0;
  uint32_t out_bytes = 64;
  std::vector<PushRegion> regions;   // what arrived, in (source, chunk) order
};

int dist_push_reserve(DistState* d, uint64_t n, uint32_t n_chunks, std::string* err);
int dist_push_fold(DistState* d, const uint8_t* d_records, uint64_t n, const PushFoldArgs& pf, cudaStream_t st, PushFoldResult* out,
                   std::string* err);
int dist_gather_regions(DistState* d, const PushFoldResult& res, const RowProgram& prog, cudaStream_t st, const uint8_t** out, std::string* err);

}  // namespace sgr
