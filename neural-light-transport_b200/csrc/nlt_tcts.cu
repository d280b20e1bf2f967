// tcgen05 "TS" building blocks: the A operand of the MMA lives in TENSOR MEMORY (written there from registers by
// tcgen05.st), B in shared memory.  For the NLT convolutions the activation operand needs a per-element fp32 ->
// (hi, lo) TF32 split before it can enter the tensor core; doing that split on the way from shared memory to TMEM
// (thread = pixel row = TMEM lane) removes the smem -> smem transform pass and two thirds of the MMA's shared-memory
// operand reads of the SS form in nlt_tc.cu.
//
// This file starts with a self-checking probe of exactly the hardware behaviour the kernels rely on (TMEM layout of
// a TF32 A operand, tcgen05.st -> MMA ordering, what the tensor core does with the low 13 mantissa bits of an
// unrounded fp32 input); tests/test_gpu_tcts.py runs it.
#include <cuda.h>
#include "nlt_common.cuh"

namespace nlt {

__device__ __forceinline__ uint32_t ts_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ts_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool ts_mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void ts_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!ts_mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) { asm volatile("trap;"); }     // watchdog: trap instead of hanging the GPU
  }
}
__device__ __forceinline__ void ts_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ts_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ts_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem], TF32 inputs, fp32 accumulate, M = 128
__device__ __forceinline__ void ts_mma_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void ts_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void ts_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ts_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major B tile with 64-byte rows (16 tf32), SWIZZLE_64B, 8-row atoms (as umma_desc_kmajor<64> in nlt_tc.cu)
__device__ __forceinline__ uint64_t ts_desc_k64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * 64) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// Probe: out[128 x BN] = A[128 x 16] * B[BN x 16]^T with A written to TMEM by tcgen05.st (thread t = row t, the 16 K
// values in 16 consecutive columns) and B staged in shared memory in the SWIZZLE_64B K-major layout.
// One CTA of 128 threads.  A and B are passed as plain row-major fp32 arrays; A is NOT rounded here (so the host can
// see what the tensor core does with the low mantissa bits).
template <int BN>
__global__ void __launch_bounds__(128, 1)
tcts_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out) {
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  __shared__ __align__(1024) float bs[BN * 16];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) {
    ts_mbar_init(ts_smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ts_smem_u32(&tmem_base_smem)),
                 "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B: row n = 64 bytes = four 16-byte chunks, physical chunk = c ^ ((n >> 1) & 3)
  for (int i = t; i < BN * 16; i += 128) {
    const int n = i >> 4, k = i & 15;
    const int chunk = (k >> 2) ^ ((n >> 1) & 3);
    bs[n * 16 + chunk * 4 + (k & 3)] = B[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the MMA
  ts_fence_before();
  __syncthreads();
  ts_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // A row of this thread -> TMEM lane (32*warp + lane), columns [32, 48)
  uint32_t a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = __float_as_uint(A[t * 16 + k]);
  ts_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + 32u, a);
  ts_wait_st();
  ts_fence_before();
  __syncthreads();
  ts_fence_after();
  if (t == 0) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      ts_mma_tf32(tmem_base, tmem_base + 32u + 8u * ks, ts_desc_k64(ts_smem_u32(bs) + 32u * ks), IDESC, ks != 0);
    ts_commit(ts_smem_u32(&bar));
  }
  ts_mbar_wait(ts_smem_u32(&bar), 0);
  ts_fence_after();
#pragma unroll
  for (int c0 = 0; c0 < BN; c0 += 16) {
    uint32_t v[16];
    ts_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) out[t * BN + c0 + j] = __uint_as_float(v[j]);
  }
  ts_fence_before();
  __syncthreads();
  if (warp == 0) {
    ts_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

}  // namespace nlt

using namespace nlt;

extern "C" {

// Diagnostic entry point (tests/test_gpu_tcts.py): out[128 x bn] = A[128 x 16] * B[bn x 16]^T through the TS-form MMA.
int nlt_debug_tcts_probe(const float* A, const float* B, int32_t bn, float* out, void* stream) {
  NLT_CHECK_ARG(A && B && out && (bn == 16 || bn == 32), "tcts_probe: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (bn == 16) tcts_probe_kernel<16><<<1, 128, 0, st>>>(A, B, out);
  else tcts_probe_kernel<32><<<1, 128, 0, st>>>(A, B, out);
  NLT_CUDA_LAUNCH_CHECK("tcts_probe_kernel");
  return NLT_OK;
}

}  // extern "C"
