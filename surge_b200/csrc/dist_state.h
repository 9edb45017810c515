// dist_state.h — state of one rank of the multi-GPU replay (shared by dist.cu and route_push.cu).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "devbuf.h"
#include "dist.cuh"

namespace sgr {

// Every rank's receive allocation starts with a header the other ranks write arrival flags into
// (u64 flags[kMaxRanks][kMaxChunks]: (epoch << 32) | records + 1 of region (source, chunk)); the records follow.
constexpr size_t kRecvHeaderBytes = 64 << 10;
constexpr int kMaxChunks = 256;

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string* err);
};
NcclApi& nccl_api();

struct DistState {
  int rank = 0, nranks = 1;
  bool loopback = false;                 // several ranks inside one process (tests): no NCCL, peers handed over as raw pointers
  ncclComm_t comm = nullptr;
  uint64_t n_global = 0, n_local = 0;
  DevBuf owner_of, local_of, global_of_local, part_tmp, flags, pos, scan_tmp;
  DevBuf hist, owner_total, counts_all, send_buf, recv_buf;
  uint64_t recv_capacity = 0;            // records
  uint8_t* peer_recv[kMaxRanks] = {};    // every rank's receive RECORDS (behind the header), mapped here
  uint8_t* peer_base[kMaxRanks] = {};    // every rank's receive allocation (header first)
  bool peers_mapped = false;
  std::vector<void*> opened;             // IPC mappings to close
  DistStats stats{};
  cudaEvent_t ev[6] = {};
  // pipelined push path (route_push.cu)
  DevBuf route_of;                       // owner << 28 | local index, per global aggregate
  DevBuf lb, push_ctl, gather_buf;       // look-back cells; tickets + chunk totals + status; contiguous copy for the replay
  cudaStream_t stream2 = nullptr, stream_hi = nullptr;   // fold (low priority) and partition (high priority) streams
  cudaEvent_t pev[4] = {};
  uint32_t epoch = 0;
  void* h_pinned = nullptr;              // page-locked landing area of the per-call read-backs
};

}  // namespace sgr
