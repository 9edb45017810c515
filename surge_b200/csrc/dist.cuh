// dist.cuh — multi-GPU routing (K4 + exchange) interface used by engine.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace sgr {

constexpr int kMaxRanks = 16;

struct DistStats {
  float ms_count = 0, ms_counts_exchange = 0, ms_scatter = 0, ms_exchange = 0;
  uint64_t n_sent = 0, n_sent_remote = 0, n_recv = 0;
};

struct DistState;
DistState* dist_create();
void dist_destroy(DistState* d);
int dist_unique_id(void* out128, std::string* err);
int dist_init(DistState* d, int rank, int nranks, const void* unique_id, uint64_t recv_capacity_records, cudaStream_t st, std::string* err);
int dist_ipc_export(DistState* d, void* out64, std::string* err);
int dist_ipc_import(DistState* d, const void* handles, std::string* err);
int dist_set_peers(DistState* d, void* const* bases, std::string* err);
void* dist_recv_base(const DistState* d);
int dist_set_partitions(DistState* d, const uint32_t* partition_of_agg, uint64_t n_global, cudaStream_t st, std::string* err);
int dist_route(DistState* d, const uint8_t* d_records, uint64_t n, bool fused, unsigned long long* d_counters, cudaStream_t st,
               uint64_t* n_recv_out, std::string* err);
uint64_t dist_n_local(const DistState* d);
int dist_nranks(const DistState* d);
bool dist_is_loopback(const DistState* d);
void dist_clear_stats(DistState* d, uint64_t n_records);
const uint32_t* dist_global_of_local(const DistState* d);
const DistStats* dist_stats(const DistState* d);
const uint8_t* dist_recv_buffer(const DistState* d);

}  // namespace sgr
