// fold_kernels.cuh — launch interface of the segmented event-fold kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sgr_device.cuh"

namespace sgr {

struct FoldArgs {
  const uint8_t* events;         // event log base, 16-byte aligned
  const uint64_t* seg_offsets;   // n_seg+1 byte offsets (multiples of 16)
  const uint32_t* seg_ids;       // optional: state slot of segment i (incremental fold); null => slot i
  const uint32_t* seg_list;      // optional: fold only these segments (n_seg of them), e.g. an exact-replay list
  uint64_t n_seg;
  const unsigned long long* n_seg_dev;  // optional: min(*n_seg_dev, n_seg) segments (count produced on the device)
  const uint8_t* states_in;      // optional prior states, slot-indexed; null => all None
  uint8_t* states_out;           // slot-indexed; may alias states_in
  unsigned long long* counters;  // [0] events applied, [1] aggregates in error, [2] segments left to the split path,
                                 // [4] records dropped after a throw (fixed records), [5] events applied in replay mode
  uint64_t long_threshold;       // segments longer than this many bytes are skipped here (0 = never)
};

struct FoldLaunchInfo {
  int variant;       // index into the config table actually used
  int threads, chunk, stages, grid;
  size_t smem;
};

// Launch the streaming fold (K1 fixed / K2 variable records). variant < 0 picks the default
// for the record kind. Returns cudaSuccess or the launch error.
cudaError_t launch_fold_stream(const FoldArgs& args, const DevProgram& prog, int variant, int num_sms,
                               uint32_t max_record_bytes, cudaStream_t stream, FoldLaunchInfo* info);

int fold_variant_count();
const char* fold_variant_name(int variant);

}  // namespace sgr
