// incremental.cuh — K6 (sort-free micro-batch fold for programs inside the transformer algebra).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fold_rows.cuh"

namespace sgr {
size_t inc_scratch_bytes(uint64_t n_slots);
cudaError_t launch_incremental_atomic(const uint8_t* d_records, uint32_t n, uint64_t n_slots, void* d_scratch, uint8_t* d_states,
                                      uint32_t* d_touched_ids, uint32_t* d_err_ids, const uint32_t* d_prev_ids,
                                      const unsigned long long* d_prev_n, uint32_t prev_n_upper, const RowProgram& prog,
                                      unsigned long long* d_counters, unsigned long long replay_budget, cudaStream_t st);
}  // namespace sgr
