// sgr_device.cuh — device-side helpers shared by the kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "surge_b200 kernels are written for sm_100a only"
#endif

namespace sgr {

// ---------------------------------------------------------------- compact device form of sgr_fold_program
// One op word: opcode[3:0] | nwords[9:4] | dst_word[15:10] | src_word[31:16]
struct DevRule {
  uint32_t exists_rule;
  uint32_t n_ops;
  uint32_t min_len;   // max over ops of src_off+len: shortest record this rule can read
  uint32_t pad;
  uint32_t ops[8];
};
struct DevProgram {
  uint32_t state_words;  // state_bytes / 4, including the 2 engine words
  uint32_t user_words;   // state_words - 2
  uint32_t record_kind;
  uint32_t n_types;
  uint32_t n_f64;
  uint32_t f64_word[8];  // word index of the low half of each f64 state field
  uint32_t pad[3];
  DevRule rules[16];
};
static_assert(sizeof(DevRule) == 48, "DevRule layout");
static_assert(sizeof(DevProgram) % 16 == 0, "DevProgram must be copyable as uint4");

__host__ __device__ inline uint32_t pack_op(uint32_t opcode, uint32_t nwords, uint32_t dst_word, uint32_t src_word) {
  return (opcode & 15u) | ((nwords & 63u) << 4) | ((dst_word & 63u) << 10) | (src_word << 16);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------- PTX wrappers: mbarrier + 1-D TMA bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.
// dst, src 16-byte aligned; bytes a non-zero multiple of 16. SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
#endif  // __CUDACC__

}  // namespace sgr
