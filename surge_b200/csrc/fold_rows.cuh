// fold_rows.cuh — launch interface of the record-parallel fold (K1/K3, fixed 64-byte records).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "sgr_device.cuh"

namespace sgr {

constexpr int kRowThreads = 256;
constexpr int kMaxSlots = 16;   // distinct record words a program may read (slot 0 = event type)
constexpr int kMaxRowWords = 14;  // widest state in transformer form: 64-byte struct
constexpr int kTabStride = 16;   // words per type in RowProgram::tab

// Program in transformer form: per event type, how each state word is produced.
// Two closed classes of programs (build_row_program decides):
//   class 0  MATERIALISE / CREATE / TOMBSTONE / THROW rules          (Counter, IntBalance, snapshot restore)
//   class 1  IF_EXISTS / CREATE / TOMBSTONE / THROW rules            (BankAccount): an IF_EXISTS event applies iff the
//            state exists at that point, i.e. iff the input existed or a CREATE came before it — one extra
//            composition rule (a tombstoned prefix absorbs it), still associative
// A program mixing MATERIALISE and IF_EXISTS is outside both and takes the lane-sequential kernel.
struct RowProgram {
  uint32_t user_words;
  uint32_t n_slots;
  uint32_t cls;                   // 0 / 1, see above
  uint32_t f64_mask;              // bit w: state words w, w+1 form a JVM Double (numeric == for the publish rule)
  uint32_t slot_word[kMaxSlots];  // record word index of each slot
  uint32_t tab[16 * kTabStride];  // per type: [0] bit0 valid, bit1 result is None, bit2 IF_EXISTS rule; [1+w] mode | neg<<2 | slot<<3
};

struct RowArgs {
  const uint8_t* events;        // device log; records at log_begin + 64*i
  const uint64_t* seg_offsets;  // n_seg+1 byte offsets, all == log_begin (mod 64)
  const uint32_t* seg_ids;      // optional state slot per segment
  uint64_t n_seg;
  uint64_t log_begin, log_end;  // seg_offsets[0], seg_offsets[n_seg]
  const uint8_t* states_in;     // optional prior states
  uint8_t* states_out;
  unsigned long long* counters; // [0] events applied, [1] aggregates in error, [3] segments queued for exact replay,
                                // [4] records dropped after a throw, [6] grid-barrier arrivals (fold_runs)
  unsigned long long* counters_next;  // fold_runs: the other counter block, zeroed for the next fold
  uint32_t* redo_ids;           // segments whose handler threw (replayed by the sequential kernel)
  uint64_t redo_cap;
  uint32_t* part_flags;         // per warp: == epoch once its open transformer is published
  uint32_t* part_data;          // per warp: m, v[W], ex | has_head<<2
  uint32_t epoch;
};

// false if the program is outside the transformer algebra (IF_EXISTS rules, 64-bit adds, f64 fields,
// unsupported state width): the caller then uses the lane-sequential kernel.
bool build_row_program(const DevProgram& dp, RowProgram* out);
// one pass over the CSR offsets at load time: are all segments 64-byte aligned relative to the first, and
// where does the log begin/end (device offsets are opaque to the host otherwise)
cudaError_t inspect_offsets(const uint64_t* d_off, uint64_t n_seg, unsigned long long* d_scratch, cudaStream_t st,
                            bool* aligned64, uint64_t* log_begin, uint64_t* log_end, uint64_t* max_seg_bytes);
int row_kernel_max_grid(int num_sms, const RowProgram& prog);  // largest co-resident grid (look-back needs forward progress)
cudaError_t launch_fold_rows(const RowArgs& args, const RowProgram& prog, int grid, cudaStream_t stream);


// ---- fold_vruns.cu: variable records with a record directory
struct VarArgs {
  const uint8_t* events;
  const uint64_t* rec_offsets;   // n_rec+1 byte offsets, log order
  uint64_t n_rec;
  const uint64_t* seg_offsets;   // CSR, the source of truth (cross-checked at every segment head)
  uint64_t n_seg;
  uint8_t* states_out;           // zeroed before the launch (full rebuild: empty aggregates stay None)
  unsigned long long* counters;  // [0] records seen, [3] replay list length, [4] records of replayed segments, [7] CSR mismatches
  uint32_t* redo_ids; uint64_t redo_cap;
  uint32_t* part_flags; uint32_t* part_data; uint32_t epoch;   // part_data: 8 words per warp
  uint32_t stage_bytes;
};
int vruns_config(int num_sms, uint32_t max_record_bytes, uint32_t stage_hint, int nstage, int* threads, size_t* smem, uint32_t* stage_bytes);  // returns max grid, 0 if impossible
cudaError_t launch_fold_vruns(const VarArgs& args, const RowProgram& prog, int nstage, int grid, int threads, size_t smem, cudaStream_t stream);

// ---- fold_runs.cu: lane-run variant (primary). Same RowArgs / RowProgram.
int run_variant_count();
const char* run_variant_name(int v);
int run_kernel_max_grid(int num_sms, int variant, const RowProgram& prog);
int run_variant_step_bytes(int variant, const RowProgram& prog);
int run_warps_per_cta();
cudaError_t launch_fold_runs(const RowArgs& args, const RowProgram& prog, int variant, int grid, cudaStream_t stream);

}  // namespace sgr
