// tcgen05 tensor-core path of the generalised convolution (sm_100a only).
//
// Implicit GEMM  D[128 pixels x BN] += A[128 pixels x 32 ch] * B[32 ch x BN]  per K-block, where a
// K-block is (tap, segment, 32-channel chunk).  No im2col: for every tap the A tile is ONE TMA box of the
// NHWC activation tensor, [TH x TW pixels] x [32 channels], landing in shared memory as 128 rows of 128 B with
// the 128-byte swizzle -- exactly the K-major operand layout tcgen05.mma consumes.  SAME padding is TMA
// out-of-bounds zero fill (coordinates may be negative); the k == stride "patch" conv is a 5-D view of the
// same tensor (c, dx, x/s, dy, n*H/s + y/s) so that no element stride is needed.
//
// fp32-class accuracy on TF32 tensor cores: 3xTF32.  Weights are split once per call into hi/lo TF32
// planes (pack kernel); activations are split in shared memory by four transform warps
// (hi = rna_tf32(a) in place, lo = rna_tf32(a - hi) in a second buffer) and the MMA warp issues
// D += Ahi*Bhi + Ahi*Blo + Alo*Bhi with fp32 accumulation in TMEM.
//
// Warp roles (320 threads, 1 CTA / SM, persistent over output tiles):
//   warp 0      TMA producer      (A raw tile + Bhi + Blo per stage, mbarrier complete_tx)
//   warp 1      MMA issuer        (single thread; tcgen05.commit frees stages / publishes accumulators)
//   warps 2-5   operand transform (smem -> smem split, fence.proxy.async)
//   warps 6-9   epilogue          (tcgen05.ld -> bias / activation / beta / derivative mask -> global)
// Accumulators are double-buffered in TMEM so the epilogue of tile t overlaps the MMAs of tile t+1.
#include <cuda.h>
#include "nlt_common.cuh"

namespace nlt {

constexpr int TC_MAX_STAGES = 8;                // smem ring depth is chosen per shape (2..8): bytes in flight, not math, bound most layers
constexpr int TC_BM = 128;
// channels per K-block: 32 (128-byte rows, SWIZZLE_128B) or 16 (64-byte rows, SWIZZLE_64B)
constexpr int TC_EPI_PAD = 36;                  // floats per staging row (32 columns + 4 pad: conflict-free)
constexpr int TC_EPI_BYTES = 4 * 32 * TC_EPI_PAD * 4;   // epilogue staging: 4 warps x 32 rows
constexpr int TC_THREADS = 320;               // wgrad kernel: TMA + MMA + 8 transform warps (4 of them double as epilogue)
constexpr int TCF_THREADS = 448;              // forward kernel: TMA + MMA + 4 transform + 4 epilogue + 4 more transform warps
constexpr uint32_t TC_SPIN_LIMIT = 1u << 28;    // watchdog: trap instead of hanging the GPU
unsigned long long g_tc_launches = 0;           // tensor-core kernel launches (diagnostic)

int g_opt_tc_rawhi = -1;

struct TcParams {
  // lattice / tiling
  int N, Hl, Wl;            // lattice size (pixels this kernel enumerates)
  int TW, TH;               // tile = TH x TW lattice pixels (TW*TH == 128)
  int tiles_x, tiles_y;     // per image
  int n_tiles_n;            // column tiles (Cout / BN)
  int total_tiles;          // N * tiles_y * tiles_x * n_tiles_n
  // K-blocks
  int ntap_y, ntap_x;
  int nseg;
  int seg_chunks[NLT_MAX_SEG];
  int kb_total;
  // TMA coordinate maps: x = tx0*mx + ux*ux_step + x_off (mode A); mode B (patch) uses (ux, tx0, uy, n*Hs + ty0)
  int mode_patch;           // 1: 5-D patch view
  int ux_step, x_off, uy_step, y_off;
  int Hs;                   // mode B: H / s  (rows of the merged n*H/s dimension per image)
  // output
  int Hout, Wout, cout_true, Cout;   // Cout = GEMM columns (d2s: s*s*cout_true)
  int o0y, osy, o0x, osx;            // output pixel = o0 + os * lattice coordinate (non-d2s)
  int d2s, d2s_s;
  // exact multiply-shift division by the run-time tile / channel counts: the per-tile coordinate decode of the producer
  // and the per-pass address arithmetic of the epilogue warps were chains of 32-bit integer divisions (~150 cycles
  // each); with everything else ablated a tile still cost 4 us (profiles/r2_z_*)
  FastDiv div_ntn, div_tpi, div_tx, div_tw, div_ct, div_s;
  int act, mask_act;
  int stages;               // smem ring depth
  int rawhi;                // option "tc_rawhi" / NLT_TC_RAWHI (default 1): A-hi operand = the raw fp32 tile (no hi plane write)
  int ablate;               // DIAGNOSTIC (NLT_TC_ABLATE, wrong results!): 1 no B loads, 2 no transform, 4 no MMA, 8 no epilogue memory traffic, 16 no A loads
  float beta;
  const float* bias;
  const float* mask_y;
  float* out;
};

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// DIAGNOSTIC switch (NLT_TC_ABLATE bit 32): poll with the non-suspending test_wait instead of try_wait
__device__ int g_tc_poll = 0;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  if (g_tc_poll) {
    while (!mbar_test_wait(bar, parity)) {
      if (++spins > TC_SPIN_LIMIT) { asm volatile("trap;"); }
    }
    return;
  }
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > TC_SPIN_LIMIT) { asm volatile("trap;"); }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, fp32 accumulate, M=128
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// TS form: A operand in tensor memory (see nlt_tcts.cu for the probe of the layout this relies on)
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// Round-to-nearest TF32 in two integer instructions (cvt.rna.tf32 expands to ~8 SASS instructions with its
// inf/nan handling; ncu showed the operand split bound by exactly that).  Adding half an ulp to the bit
// pattern and clearing the low 13 mantissa bits rounds the magnitude to nearest (ties away), carries into
// the exponent correctly, and is only wrong for inf/nan, which the split does not need to preserve.
__device__ __forceinline__ float tf32_round(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// K-major swizzled operand tile: rows of ROWB bytes (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B), 8-row atoms.
template <int ROWB>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major) = 16 B
  d |= (uint64_t)((8 * ROWB) >> 4) << 32;             // stride byte offset: 8 rows
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)(ROWB == 128 ? 2 : 4) << 61;         // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

// ---------------------------------------------------------------------------------------------
// weight pack: Bhi/Blo [kb][Cout_pad][32] (K-major rows of 128 B), TF32-rounded split
// ---------------------------------------------------------------------------------------------
__global__ void tc_pack_weights_kernel(const GConvK g, int kbw, int kb_total, int chunks_per_tap, int cout_pad,
                                       float* __restrict__ bhi, float* __restrict__ blo) {
  const size_t total = (size_t)kb_total * cout_pad * kbw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % kbw);
    const int n = (int)((i / kbw) % cout_pad);
    const int kb = (int)(i / ((size_t)kbw * cout_pad));
    const int tapi = kb / chunks_per_tap;
    int ch = kb - tapi * chunks_per_tap;
    int s = 0;
    while (s < g.nseg - 1 && ch >= g.seg[s].C / kbw) { ch -= g.seg[s].C / kbw; ++s; }
    const int c = g.seg[s].coff + ch * kbw + kk;
    const int uy = tapi / g.ax.nu, ux = tapi - uy * g.ax.nu;
    float v = 0.f;
    if (n < g.Cout) {
      int tap = (g.ay.d0 + g.ay.ds * uy) * g.kw + (g.ax.d0 + g.ax.ds * ux), nn = n;
      if (g.d2s) { tap = n / g.cout_true; nn = n - tap * g.cout_true; }
      v = __ldg(g.w + (long long)tap * g.wt + (long long)c * g.wc + (long long)nn * g.wn);
    }
    const float hi = tf32_rna(v);
    bhi[i] = hi;
    blo[i] = tf32_rna(v - hi);
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
struct TcMaps {
  CUtensorMap a[NLT_MAX_SEG];
  CUtensorMap bhi, blo;
};

template <int BN, int KBW>
__global__ void __launch_bounds__(TCF_THREADS, 1)
tc_gconv_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  constexpr int ROWB = KBW * 4;                   // bytes per operand row
  constexpr int TC_KB = KBW;
  constexpr int TC_A_BYTES = TC_BM * ROWB;
  constexpr int B_BYTES = BN * ROWB;
  constexpr int STAGE_BYTES = 2 * TC_A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle
  // offset arithmetic on the __shared__ array keeps the address space (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bars[3 * TC_MAX_STAGES + 4];
  const int TC_STAGES = p.stages;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_ready = [&](int s) { return bar0 + 8u * (TC_MAX_STAGES + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (2 * TC_MAX_STAGES + s); };
  auto bar_accf = [&](int b) { return bar0 + 8u * (3 * TC_MAX_STAGES + b); };
  auto bar_acce = [&](int b) { return bar0 + 8u * (3 * TC_MAX_STAGES + 2 + b); };
  const uint32_t smem_base = smem_u32(smem);
  auto a_hi = [&](int s) { return smem_base + (uint32_t)s * STAGE_BYTES; };
  auto a_lo = [&](int s) { return smem_base + (uint32_t)s * STAGE_BYTES + TC_A_BYTES; };
  auto b_hi = [&](int s) { return smem_base + (uint32_t)s * STAGE_BYTES + 2 * TC_A_BYTES; };
  auto b_lo = [&](int s) { return smem_base + (uint32_t)s * STAGE_BYTES + 2 * TC_A_BYTES + B_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_ready(s), 8);                  // one arrival per transform warp (see the transform role)
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accf(b), 1);
      mbar_init(bar_acce(b), 4);                   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int kb_total = p.kb_total;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int mt = (int)fdiv((uint32_t)tile, p.div_ntn);
        const int nt = tile - mt * p.n_tiles_n;
        const int n = (int)fdiv((uint32_t)mt, p.div_tpi);
        const int r = mt - n * tiles_per_img;
        const int rq = (int)fdiv((uint32_t)r, p.div_tx);
        const int ty0 = rq * p.TH, tx0 = (r - rq * p.tiles_x) * p.TW;
        int kb = 0;
        for (int uy = 0; uy < p.ntap_y; ++uy)
          for (int ux = 0; ux < p.ntap_x; ++ux)
            for (int s = 0; s < p.nseg; ++s)
              for (int ch = 0; ch < p.seg_chunks[s]; ++ch, ++kb) {
                mbar_wait(bar_empty(stage), phase ^ 1);
                mbar_expect_tx(bar_full(stage), ((p.ablate & 16) ? 0 : TC_A_BYTES) + ((p.ablate & 1) ? 0 : 2 * B_BYTES));
                if (!(p.ablate & 16)) {
                  if (p.mode_patch)
                    tma_load_5d(a_hi(stage), &maps.a[s], bar_full(stage), ch * TC_KB, ux, tx0, uy, n * p.Hs + ty0);
                  else
                    tma_load_4d(a_hi(stage), &maps.a[s], bar_full(stage), ch * TC_KB, tx0 + ux * p.ux_step + p.x_off,
                                ty0 + uy * p.uy_step + p.y_off, n);
                }
                if (!(p.ablate & 1)) {
                  tma_load_3d(b_hi(stage), &maps.bhi, bar_full(stage), 0, nt * BN, kb);
                  tma_load_3d(b_lo(stage), &maps.blo, bar_full(stage), 0, nt * BN, kb);
                }
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
              }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1;
        mbar_wait(bar_acce(buf), acc_phase ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(bar_ready(stage), phase);        // operands split and visible to the async proxy
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < ((p.ablate & 4) ? 0 : TC_KB / 8); ++k) {
            const uint64_t ah = umma_desc_kmajor<ROWB>(a_hi(stage) + k * 32), al = umma_desc_kmajor<ROWB>(a_lo(stage) + k * 32);
            const uint64_t bh = umma_desc_kmajor<ROWB>(b_hi(stage) + k * 32), bl = umma_desc_kmajor<ROWB>(b_lo(stage) + k * 32);
            tc_mma_tf32(d_tmem, al, bh, IDESC, (kb | k) != 0);   // small terms first
            tc_mma_tf32(d_tmem, ah, bl, IDESC, 1);
            tc_mma_tf32(d_tmem, ah, bh, IDESC, 1);
          }
          tc_commit(bar_empty(stage));               // stage reusable once these MMAs have read it
          if (kb == kb_total - 1) tc_commit(bar_accf(buf));
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 6 || warp >= 10) {
    // ================= operand transform: fp32 -> (hi, lo) TF32 planes (warps 2-5 and 10-13) =================
    // eight warps: with four, one warp per scheduler ran this dependent LDS -> cvt -> STS chain at ~0.3 IPC
    const int t = warp < 6 ? threadIdx.x - 64 : threadIdx.x - 320 + 128;   // 0..255
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(bar_full(stage), phase);
        float4* ah = reinterpret_cast<float4*>(smem + (size_t)stage * STAGE_BYTES);
        float4* al = reinterpret_cast<float4*>(smem + (size_t)stage * STAGE_BYTES + TC_A_BYTES);
#pragma unroll
        for (int i = 0; i < ((p.ablate & 2) ? 0 : TC_A_BYTES / 16 / 256); ++i) {
          const int q = t + i * 256;              // physical 16-byte chunk: elementwise, swizzle-agnostic
          const float4 v = ah[q];
          float4 h, l;
          if (p.rawhi) {
            // the tensor core reads the upper 19 bits of an fp32 word (measured: tests/test_gpu_tcts.py): the raw tile IS
            // the hi operand, only lo = x - trunc(x) (exact in fp32, then rounded to TF32) has to be written
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            l.x = tf32_round(v.x - h.x); l.y = tf32_round(v.y - h.y); l.z = tf32_round(v.z - h.z); l.w = tf32_round(v.w - h.w);
            al[q] = l;
          } else {
            h.x = tf32_round(v.x); h.y = tf32_round(v.y); h.z = tf32_round(v.z); h.w = tf32_round(v.w);
            l.x = tf32_round(v.x - h.x); l.y = tf32_round(v.y - h.y); l.z = tf32_round(v.z - h.z); l.w = tf32_round(v.w - h.w);
            ah[q] = h;
            al[q] = l;
          }
        }
        // every thread makes its writes visible to the async proxy; ONE arrival per warp: 256 single-thread arrivals
        // on the same mbarrier word serialise (they were the largest part of the kernel's per-k-block fixed cost)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ready(stage));
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue =================
    // TMEM -> registers (lane = pixel) -> bias/activation -> per-warp smem staging -> transposed read
    // (8 lanes = 128 contiguous bytes of one pixel) -> beta / derivative-mask -> coalesced global store.
    // The read-modify-write operands of the NEXT 32-column chunk are loaded before the current one is
    // processed, so their DRAM latency overlaps the TMEM drain.
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32)
    float* stg = reinterpret_cast<float*>(smem + (size_t)TC_STAGES * STAGE_BYTES) + quarter * 32 * TC_EPI_PAD;
    const bool rmw = ((p.beta != 0.f) || (p.mask_y != nullptr)) && !(p.ablate & 8);
    constexpr int NCH = (BN + 31) / 32;           // 32-column chunks per tile (BN == 16: one half-used chunk)
    constexpr int CW = BN < 32 ? BN : 32;         // columns per chunk
    constexpr int QPR = CW / 4;                   // float4 quads per pixel row of a chunk
    constexpr int RPI = 32 / QPR;                 // pixel rows covered per pass of the warp
    constexpr int NPASS = 32 / RPI;               // passes per chunk (== QPR)
    const int lq = lane % QPR, lr = lane / QPR;

    // Output addressing without per-pass arithmetic: offset(tile, pass i, column) = tile base + row_off[i] + col_off, where
    // row_off depends only on the lane (computed once per kernel) and col_off only on (column tile, chunk).  The
    // epilogue warps are the busiest role of this kernel on the many-tile levels (ncu source view, profiles/r2_af_*:
    // ~85 % of their samples are not waits) -- every instruction here is on the critical path of a tile.
    const int sxy = p.d2s ? p.d2s_s : 1;
    int row_off[NPASS];
#pragma unroll
    for (int i = 0; i < NPASS; ++i) {
      const int row = quarter * 32 + i * RPI + lr;
      const int rty = (int)fdiv((uint32_t)row, p.div_tw), rtx = row - rty * p.TW;
      row_off[i] = p.d2s ? (rty * sxy * p.Wout + rtx * sxy) * p.cout_true
                         : (p.osy * rty * p.Wout + p.osx * rtx) * p.cout_true;
    }
    auto tile_base = [&](int tile, int& nt) -> size_t {
      const int mt = (int)fdiv((uint32_t)tile, p.div_ntn);
      nt = tile - mt * p.n_tiles_n;
      const int n = (int)fdiv((uint32_t)mt, p.div_tpi);
      const int r = mt - n * tiles_per_img;
      const int rq = (int)fdiv((uint32_t)r, p.div_tx);
      const int ty0 = rq * p.TH, tx0 = (r - rq * p.tiles_x) * p.TW;
      const int oy = p.d2s ? ty0 * sxy : p.o0y + p.osy * ty0, ox = p.d2s ? tx0 * sxy : p.o0x + p.osx * tx0;
      return (((size_t)n * p.Hout + oy) * p.Wout + ox) * p.cout_true;
    };
    // offset of GEMM column nb inside an output pixel group (d2s: which of the s x s pixels, which channel)
    auto col_off = [&](int nb) -> int {
      if (!p.d2s) return nb;
      const int tap = (int)fdiv((uint32_t)nb, p.div_ct);
      const int dy = (int)fdiv((uint32_t)tap, p.div_s);
      return (dy * p.Wout + (tap - dy * p.d2s_s)) * p.cout_true + (nb - tap * p.cout_true);
    };
    const bool bias_vec = p.bias != nullptr && aligned16(p.bias);
    const int act = p.act;
    const float neg = act == NLT_ACT_LEAKYRELU ? 0.3f : (act == NLT_ACT_RELU ? 0.f : 1.f);
    float4 rm_old[NPASS], rm_y[NPASS];
    auto prefetch = [&](size_t base) {
#pragma unroll
      for (int i = 0; i < NPASS; ++i) {
        const size_t ob = base + (size_t)row_off[i];
        rm_old[i] = (p.beta != 0.f) ? *reinterpret_cast<const float4*>(p.out + ob) : make_float4(0.f, 0.f, 0.f, 0.f);
        rm_y[i] = (p.mask_y != nullptr) ? ld4(p.mask_y + ob) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1;
      int nt;
      const size_t tbase = tile_base(tile, nt);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN);
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        const int col0 = nt * BN + ch * 32;       // GEMM column of this chunk
        const size_t cbase = tbase + (size_t)col_off(col0 + lq * 4);
        // read-modify-write operands first: their latency overlaps the accumulator wait / TMEM drain
        if (rmw) prefetch(cbase);
        if (ch == 0) {
          mbar_wait(bar_accf(buf), acc_phase);
          tc_fence_after();
        }
#pragma unroll
        for (int h = 0; h < CW / 16; ++h) {
          uint32_t v[16];
          tc_ld16(taddr + ch * 32 + h * 16, v);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            int cb = col0 + h * 16 + q4 * 4;
            if (p.d2s) cb -= (int)fdiv((uint32_t)cb, p.div_ct) * p.cout_true;
            float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias_vec) bq = ld4(p.bias + cb);
            else if (p.bias != nullptr) bq = make_float4(__ldg(p.bias + cb), __ldg(p.bias + cb + 1), __ldg(p.bias + cb + 2), __ldg(p.bias + cb + 3));
            float o[4] = {__uint_as_float(v[q4 * 4 + 0]) + bq.x, __uint_as_float(v[q4 * 4 + 1]) + bq.y,
                          __uint_as_float(v[q4 * 4 + 2]) + bq.z, __uint_as_float(v[q4 * 4 + 3]) + bq.w};
            if (act == NLT_ACT_ELU) {             // warp-uniform
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = elu_slow(o[e]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = o[e] > 0.f ? o[e] : neg * o[e];
            }
            *reinterpret_cast<float4*>(stg + lane * TC_EPI_PAD + h * 16 + q4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NPASS; ++i) {
          const int row = i * RPI + lr;
          float4 o = *reinterpret_cast<const float4*>(stg + row * TC_EPI_PAD + lq * 4);
          if (rmw) {
            o.x += p.beta * rm_old[i].x; o.y += p.beta * rm_old[i].y; o.z += p.beta * rm_old[i].z; o.w += p.beta * rm_old[i].w;
            if (p.mask_y != nullptr) {
              o.x *= act_bwd_from_y(rm_y[i].x, p.mask_act); o.y *= act_bwd_from_y(rm_y[i].y, p.mask_act);
              o.z *= act_bwd_from_y(rm_y[i].z, p.mask_act); o.w *= act_bwd_from_y(rm_y[i].w, p.mask_act);
            }
          }
          const size_t ob = cbase + (size_t)row_off[i];
          if (!(p.ablate & 8) || o.x == 123456.f) *reinterpret_cast<float4*>(p.out + ob) = o;
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce(buf));
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, int KBW>
__global__ void __launch_bounds__(TCF_THREADS, 1)
tcs_gconv_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  constexpr int ROWB = KBW * 4;                   // bytes per operand row
  constexpr int TC_KB = KBW;
  constexpr int TC_A_BYTES = TC_BM * ROWB;
  constexpr int B_BYTES = BN * ROWB;
  constexpr int STAGE_BYTES = TC_A_BYTES + 2 * B_BYTES;          // raw A tile + Bhi + Blo (no A-lo plane: A goes to TMEM)
  constexpr int NTA = (512 - 2 * BN) / (2 * KBW) < 8 ? (512 - 2 * BN) / (2 * KBW) : 8;   // TMEM stages of the A operand
  constexpr uint32_t ACOL0 = 2 * BN;                             // first A column (behind the two accumulators)
  constexpr uint32_t TMEM_NEED = ACOL0 + NTA * 2 * KBW;
  constexpr uint32_t TMEM_COLS = TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle
  // offset arithmetic on the __shared__ array keeps the address space (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bars[4 * TC_MAX_STAGES + 4];
  const int TC_STAGES = p.stages;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };                              // TMA -> converters / MMA (smem stage)
  auto bar_empty = [&](int s) { return bar0 + 8u * (TC_MAX_STAGES + s); };           // converters (256) + MMA commit (1) -> TMA
  auto bar_aready = [&](int t) { return bar0 + 8u * (2 * TC_MAX_STAGES + t); };      // converters -> MMA (TMEM stage)
  auto bar_afree = [&](int t) { return bar0 + 8u * (3 * TC_MAX_STAGES + t); };       // MMA commit -> converters
  auto bar_accf = [&](int b) { return bar0 + 8u * (4 * TC_MAX_STAGES + b); };
  auto bar_acce = [&](int b) { return bar0 + 8u * (4 * TC_MAX_STAGES + 2 + b); };
  const uint32_t smem_base = smem_u32(smem);
  auto a_hi = [&](int s) { return smem_base + (uint32_t)s * STAGE_BYTES; };          // the raw fp32 A tile
  auto b_hi = [&](int s) { return smem_base + (uint32_t)s * STAGE_BYTES + TC_A_BYTES; };
  auto b_lo = [&](int s) { return smem_base + (uint32_t)s * STAGE_BYTES + TC_A_BYTES + B_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 257);
    }
    for (int t = 0; t < NTA; ++t) {
      mbar_init(bar_aready(t), 256);
      mbar_init(bar_afree(t), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accf(b), 1);
      mbar_init(bar_acce(b), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int kb_total = p.kb_total;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int mt = (int)fdiv((uint32_t)tile, p.div_ntn);
        const int nt = tile - mt * p.n_tiles_n;
        const int n = (int)fdiv((uint32_t)mt, p.div_tpi);
        const int r = mt - n * tiles_per_img;
        const int rq = (int)fdiv((uint32_t)r, p.div_tx);
        const int ty0 = rq * p.TH, tx0 = (r - rq * p.tiles_x) * p.TW;
        int kb = 0;
        for (int uy = 0; uy < p.ntap_y; ++uy)
          for (int ux = 0; ux < p.ntap_x; ++ux)
            for (int s = 0; s < p.nseg; ++s)
              for (int ch = 0; ch < p.seg_chunks[s]; ++ch, ++kb) {
                mbar_wait(bar_empty(stage), phase ^ 1);
                mbar_expect_tx(bar_full(stage), TC_A_BYTES + 2 * B_BYTES);   // raw A + Bhi + Blo
                if (p.mode_patch)
                  tma_load_5d(a_hi(stage), &maps.a[s], bar_full(stage), ch * TC_KB, ux, tx0, uy, n * p.Hs + ty0);
                else
                  tma_load_4d(a_hi(stage), &maps.a[s], bar_full(stage), ch * TC_KB, tx0 + ux * p.ux_step + p.x_off,
                              ty0 + uy * p.uy_step + p.y_off, n);
                tma_load_3d(b_hi(stage), &maps.bhi, bar_full(stage), 0, nt * BN, kb);
                tma_load_3d(b_lo(stage), &maps.blo, bar_full(stage), 0, nt * BN, kb);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
              }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0, ts = 0;
      uint32_t phase = 0, tphase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1;
        mbar_wait(bar_acce(buf), acc_phase ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(bar_full(stage), phase);         // B planes of this stage have landed
          mbar_wait(bar_aready(ts), tphase);         // A (hi, lo) of this k-block is in tensor memory
          tc_fence_after();
          const uint32_t a_col = tmem_base + ACOL0 + (uint32_t)(ts * 2 * KBW);
#pragma unroll
          for (int k = 0; k < TC_KB / 8; ++k) {
            const uint64_t bh = umma_desc_kmajor<ROWB>(b_hi(stage) + k * 32), bl = umma_desc_kmajor<ROWB>(b_lo(stage) + k * 32);
            tc_mma_tf32_ts(d_tmem, a_col + KBW + 8 * k, bh, IDESC, (kb | k) != 0);   // Alo * Bhi   (small terms first)
            tc_mma_tf32_ts(d_tmem, a_col + 8 * k, bl, IDESC, 1);                     // Ahi * Blo
            tc_mma_tf32_ts(d_tmem, a_col + 8 * k, bh, IDESC, 1);                     // Ahi * Bhi
          }
          tc_commit(bar_empty(stage));               // B planes of the smem stage are free once these MMAs have read them
          tc_commit(bar_afree(ts));                  // ... and so is the TMEM stage
          if (kb == kb_total - 1) tc_commit(bar_accf(buf));
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          if (++ts == NTA) { ts = 0; tphase ^= 1; }
        }
      }
    }
  } else if (warp < 6 || warp >= 10) {
    // ================= operand conversion: raw fp32 rows in shared memory -> (hi, lo) TF32 columns in TMEM =================
    // thread = pixel row = TMEM lane (a warp may only touch the lane quarter warp % 4); two warps per quarter share
    // the KBW channels of a k-block.  One conflict-free LDS.128 per four channels (the TMA swizzle is undone in the
    // address), the split in registers, tcgen05.st -- no smem -> smem pass, no A-lo plane, no proxy fence.
    const int quarter = warp & 3, half = warp >= 10 ? 1 : 0;
    const int r = quarter * 32 + lane;
    constexpr int CH = KBW / 2;                      // channels per thread and k-block
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int stage = 0, ts = 0;
    uint32_t phase = 0, tphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(bar_full(stage), phase);
        const uint8_t* row = smem + (size_t)stage * STAGE_BYTES + (size_t)r * ROWB;
        float4 v[CH / 4];
#pragma unroll
        for (int i = 0; i < CH / 4; ++i) {
          const int c = half * (CH / 4) + i;        // logical 16-byte chunk of the row
          const int pc = ROWB == 128 ? (c ^ (r & 7)) : (c ^ ((r >> 1) & 3));
          v[i] = *reinterpret_cast<const float4*>(row + pc * 16);
        }
        uint32_t hi[CH], lo[CH];
#pragma unroll
        for (int i = 0; i < CH / 4; ++i) {
          const float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // the tensor core TRUNCATES a 32-bit TF32 operand to its upper 19 bits (measured: tests/test_gpu_tcts.py,
            // profiles/r2_c_tcts_probe.json), so the raw fp32 word IS the hi part; lo = x - trunc(x) is exact in
            // fp32 and is rounded to TF32 so that the dominant residual stays unbiased
            const uint32_t xb = __float_as_uint(x[e]);
            hi[4 * i + e] = xb;
            lo[4 * i + e] = __float_as_uint(tf32_round(x[e] - __uint_as_float(xb & 0xFFFFE000u)));
          }
        }
        mbar_wait(bar_afree(ts), tphase ^ 1);        // the MMAs that read this TMEM stage last time are done
        tc_fence_after();
        const uint32_t a_col = lane_base + ACOL0 + (uint32_t)(ts * 2 * KBW) + (uint32_t)(half * CH);
        if constexpr (CH == 16) { tc_st16(a_col, hi); tc_st16(a_col + KBW, lo); }
        else { tc_st8(a_col, hi); tc_st8(a_col + KBW, lo); }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar_aready(ts));
        mbar_arrive(bar_empty(stage));               // this thread is done reading the raw tile
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        if (++ts == NTA) { ts = 0; tphase ^= 1; }
      }
    }
  } else {
    // ================= epilogue =================
    // TMEM -> registers (lane = pixel) -> bias/activation -> per-warp smem staging -> transposed read
    // (8 lanes = 128 contiguous bytes of one pixel) -> beta / derivative-mask -> coalesced global store.
    // The read-modify-write operands of the NEXT 32-column chunk are loaded before the current one is
    // processed, so their DRAM latency overlaps the TMEM drain.
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32)
    float* stg = reinterpret_cast<float*>(smem + (size_t)TC_STAGES * STAGE_BYTES) + quarter * 32 * TC_EPI_PAD;
    const bool rmw = (p.beta != 0.f) || (p.mask_y != nullptr);
    constexpr int NCH = (BN + 31) / 32;           // 32-column chunks per tile (BN == 16: one half-used chunk)
    constexpr int CW = BN < 32 ? BN : 32;         // columns per chunk
    constexpr int QPR = CW / 4;                   // float4 quads per pixel row of a chunk
    constexpr int RPI = 32 / QPR;                 // pixel rows covered per pass of the warp
    constexpr int NPASS = 32 / RPI;               // passes per chunk (== QPR)
    const int lq = lane % QPR, lr = lane / QPR;

    struct Coord { int n, ty0, tx0, nt; };
    auto tile_coord = [&](int tile) {
      Coord c;
      const int mt = (int)fdiv((uint32_t)tile, p.div_ntn);
      c.nt = tile - mt * p.n_tiles_n;
      c.n = (int)fdiv((uint32_t)mt, p.div_tpi);
      const int r = mt - c.n * tiles_per_img;
      const int rq = (int)fdiv((uint32_t)r, p.div_tx);
      c.ty0 = rq * p.TH; c.tx0 = (r - rq * p.tiles_x) * p.TW;
      return c;
    };
    // element offset of (tile row, GEMM column) in the output tensor
    auto out_off = [&](const Coord& c, int row, int nb) -> size_t {
      const int rty = (int)fdiv((uint32_t)row, p.div_tw);
      const int ty = c.ty0 + rty, tx = c.tx0 + row - rty * p.TW;
      int cb = nb, oy, ox;
      if (p.d2s) {
        const int tap = (int)fdiv((uint32_t)nb, p.div_ct);
        cb = nb - tap * p.cout_true;
        const int dy = (int)fdiv((uint32_t)tap, p.div_s);
        oy = ty * p.d2s_s + dy; ox = tx * p.d2s_s + (tap - dy * p.d2s_s);
      } else {
        oy = p.o0y + p.osy * ty; ox = p.o0x + p.osx * tx;
      }
      return (((size_t)c.n * p.Hout + oy) * p.Wout + ox) * p.cout_true + cb;
    };
    float4 rm_old[NPASS], rm_y[NPASS];
    auto prefetch = [&](const Coord& c, int ch) {
#pragma unroll
      for (int i = 0; i < NPASS; ++i) {
        const size_t ob = out_off(c, quarter * 32 + i * RPI + lr, c.nt * BN + ch * 32 + lq * 4);
        rm_old[i] = (p.beta != 0.f) ? *reinterpret_cast<const float4*>(p.out + ob) : make_float4(0.f, 0.f, 0.f, 0.f);
        rm_y[i] = (p.mask_y != nullptr) ? ld4(p.mask_y + ob) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1;
      const Coord tc = tile_coord(tile);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN);
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        // read-modify-write operands first: their latency overlaps the accumulator wait / TMEM drain
        if (rmw) prefetch(tc, ch);
        if (ch == 0) {
          mbar_wait(bar_accf(buf), acc_phase);
          tc_fence_after();
        }
        const int col0 = tc.nt * BN + ch * 32;    // GEMM column of this chunk
#pragma unroll
        for (int h = 0; h < CW / 16; ++h) {
          uint32_t v[16];
          tc_ld16(taddr + ch * 32 + h * 16, v);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            int cb = col0 + h * 16 + q4 * 4;
            if (p.d2s) cb -= (int)fdiv((uint32_t)cb, p.div_ct) * p.cout_true;
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = __uint_as_float(v[q4 * 4 + e]);
              if (p.bias != nullptr) x += __ldg(p.bias + cb + e);
              o[e] = act_fwd(x, p.act);
            }
            *reinterpret_cast<float4*>(stg + lane * TC_EPI_PAD + h * 16 + q4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NPASS; ++i) {
          const int row = i * RPI + lr;
          float4 o = *reinterpret_cast<const float4*>(stg + row * TC_EPI_PAD + lq * 4);
          if (rmw) {
            o.x += p.beta * rm_old[i].x; o.y += p.beta * rm_old[i].y; o.z += p.beta * rm_old[i].z; o.w += p.beta * rm_old[i].w;
            if (p.mask_y != nullptr) {
              o.x *= act_bwd_from_y(rm_y[i].x, p.mask_act); o.y *= act_bwd_from_y(rm_y[i].y, p.mask_act);
              o.z *= act_bwd_from_y(rm_y[i].z, p.mask_act); o.w *= act_bwd_from_y(rm_y[i].w, p.mask_act);
            }
          }
          const size_t ob = out_off(tc, quarter * 32 + row, col0 + lq * 4);
          *reinterpret_cast<float4*>(p.out + ob) = o;
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(bar_acce(buf));
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

struct TcPlan {
  bool ok;
  bool ts;      // TS form (tcs_gconv_kernel)
  int bn;
  int kbw;
  TcParams p;
  int chunks_per_tap;
  int cout_pad;
  size_t pack_floats;   // per plane
  size_t smem_bytes;
};

static int pick_bn(int cout) {
  if (cout % 128 == 0) return 128;
  if (cout % 64 == 0) return 64;
  if (cout % 32 == 0) return 32;
  if (cout % 16 == 0) return 16;
  return 0;
}

// "tcs" (TS form: A operand converted on the way into tensor memory, tcs_gconv_kernel): option / NLT_TCS,
// 1 = use it wherever the tensor path applies (and from NLT_TCS_KMIN contraction terms on), 0 = SS form only
int g_opt_tcs = -1;
static int g_tcs_kmin = -1;
static bool tcs_enabled() {
  if (g_opt_tcs < 0) { const char* e = getenv("NLT_TCS"); g_opt_tcs = e ? (atoi(e) != 0) : NLT_TCS_DEFAULT; }
  if (g_tcs_kmin < 0) { const char* e = getenv("NLT_TCS_KMIN"); g_tcs_kmin = e ? atoi(e) : 64; }
  return g_opt_tcs == 1;
}

// Which single-phase GConvK shapes the tensor path takes (everything else stays on the SIMT kernels).
static TcPlan tc_plan(const GConvK& k) {
  TcPlan pl;
  memset(&pl, 0, sizeof(pl));
  pl.ok = false;
  const bool ts = tcs_enabled();
  pl.ts = ts;
  if (k.cout_true % 4 != 0) return pl;
  pl.bn = pick_bn(k.Cout);
  if (pl.bn == 0) return pl;
  int ctot = 0;
  pl.kbw = 32;
  for (int s = 0; s < k.nseg; ++s) {
    const Seg& sg = k.seg[s];
    if (sg.C % 16 != 0 || sg.sub != nullptr || sg.bcast || !aligned16(sg.ptr)) return pl;
    if (sg.C % 32 != 0) pl.kbw = 16;
    ctot += sg.C;
  }
  const int TC_KB = pl.kbw;
  if (!aligned16(k.w)) { /* weights are only read by the pack kernel: no alignment needed */ }
  if (ts) {
    if (k.ay.nu * k.ax.nu * ctot < g_tcs_kmin) return pl;
  } else {
    // with less than 64 contraction terms the op is a pure stream: the pointwise / fp32 kernels are faster there
    if (k.ay.nu * k.ax.nu * ctot < 64) return pl;
    // measured: 16-channel (64-byte-row) K-blocks only pay off from K = 128 up
    if (pl.kbw == 16 && k.ay.nu * k.ax.nu * ctot < 128) return pl;
  }
  if (k.ay.nu < 1 || k.ax.nu < 1) return pl;
  TcParams& p = pl.p;
  p.N = k.N; p.Hl = k.ay.nt; p.Wl = k.ax.nt;
  // tile shape
  if (p.Wl >= TC_BM) {
    if (p.Wl % TC_BM) return pl;
    p.TW = TC_BM; p.TH = 1;
  } else {
    if (TC_BM % p.Wl) return pl;
    p.TW = p.Wl; p.TH = TC_BM / p.Wl;
    if (p.Hl % p.TH) return pl;
  }
  p.tiles_x = p.Wl / p.TW; p.tiles_y = p.Hl / p.TH;
  // coordinate map
  const bool stride1 = (k.ay.it == 1 && k.ax.it == 1);
  const bool patch = (!k.d2s && k.ay.it > 1 && k.ay.it == k.ax.it && k.ay.iu == 1 && k.ax.iu == 1 && k.ay.i0 == 0 &&
                      k.ax.i0 == 0 && k.ay.nu == k.ay.it && k.ax.nu == k.ax.it && k.Hin == k.ay.nt * k.ay.it &&
                      k.Win == k.ax.nt * k.ax.it);
  if (!stride1 && !patch) return pl;
  p.mode_patch = patch ? 1 : 0;
  p.ux_step = k.ax.iu; p.x_off = k.ax.i0; p.uy_step = k.ay.iu; p.y_off = k.ay.i0;
  p.Hs = patch ? k.Hin / k.ay.it : 0;
  if (patch && (k.ay.it > 256 || p.TW > 256)) return pl;
  p.ntap_y = k.ay.nu; p.ntap_x = k.ax.nu;
  p.nseg = k.nseg;
  pl.chunks_per_tap = 0;
  for (int s = 0; s < k.nseg; ++s) { p.seg_chunks[s] = k.seg[s].C / TC_KB; pl.chunks_per_tap += p.seg_chunks[s]; }
  p.kb_total = p.ntap_y * p.ntap_x * pl.chunks_per_tap;
  p.Hout = k.Hout; p.Wout = k.Wout; p.cout_true = k.cout_true; p.Cout = k.Cout;
  p.o0y = k.ay.o0; p.osy = k.ay.os; p.o0x = k.ax.o0; p.osx = k.ax.os;
  p.d2s = k.d2s; p.d2s_s = k.d2s_s;
  // few output tiles: narrower column tiles so that more SMs get work (A is re-read from L2)
  // (only while the finer tiling still fits ONE wave of 148 CTAs: every tile pays the ~10 us pipeline fill / drain of
  // the kernel again, profiles/r2_i_tc_ablation_31.txt; NLT_TC_SPLIT_WAVES=1 restores the round-1 rule "< 148 tiles")
  static int split_waves = -1;
  if (split_waves < 0) { const char* e = getenv("NLT_TC_SPLIT_WAVES"); split_waves = (e && e[0] == '1') ? 1 : 0; }
  while (pl.bn > 32) {
    const long long tiles = (long long)p.N * p.tiles_x * p.tiles_y * (k.Cout / pl.bn);
    if (split_waves ? tiles >= 148 : 2 * tiles > 148) break;
    pl.bn /= 2;
  }
  p.n_tiles_n = k.Cout / pl.bn;
  p.div_ntn = make_fastdiv((uint32_t)p.n_tiles_n); p.div_tpi = make_fastdiv((uint32_t)(p.tiles_x * p.tiles_y));
  p.div_tx = make_fastdiv((uint32_t)p.tiles_x); p.div_tw = make_fastdiv((uint32_t)p.TW);
  p.div_ct = make_fastdiv((uint32_t)p.cout_true); p.div_s = make_fastdiv((uint32_t)(p.d2s ? p.d2s_s : 1));
  const long long tt = (long long)p.N * p.tiles_x * p.tiles_y * p.n_tiles_n;
  if (tt > (1ll << 30)) return pl;
  p.total_tiles = (int)tt;
  pl.cout_pad = k.Cout;
  pl.pack_floats = (size_t)p.kb_total * pl.cout_pad * TC_KB;
  {
    const size_t stage_bytes = (ts ? 1 : 2) * (size_t)TC_BM * TC_KB * 4 + 2 * (size_t)pl.bn * TC_KB * 4;
    int st = (int)((227 * 1024 - TC_EPI_BYTES - 2048) / stage_bytes);
    if (st > TC_MAX_STAGES) st = TC_MAX_STAGES;
    if (st < 2) return pl;
    p.stages = st;
    pl.smem_bytes = (size_t)st * stage_bytes + TC_EPI_BYTES + 1024;
  }
  if (get_encode() == nullptr) return pl;
  {
    static int ablate = -1;
    if (ablate < 0) {
      const char* e = getenv("NLT_TC_ABLATE");
      ablate = e ? atoi(e) : 0;
      if (ablate & 32) { const int one = 1; cudaMemcpyToSymbol(g_tc_poll, &one, sizeof(int)); }
    }
    p.ablate = ablate;
    if (g_opt_tc_rawhi < 0) { const char* e = getenv("NLT_TC_RAWHI"); g_opt_tc_rawhi = (e && e[0] == '0') ? 0 : 1; }
    p.rawhi = g_opt_tc_rawhi;
  }
  pl.ok = true;
  return pl;
}

static int encode_maps(const GConvK& k, const TcPlan& pl, const float* bhi, const float* blo, TcMaps* maps) {
  EncodeTiledFn enc = get_encode();
  const TcParams& p = pl.p;
  const int TC_KB = pl.kbw;
  const CUtensorMapSwizzle swz = pl.kbw == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  for (int s = 0; s < k.nseg; ++s) {
    const Seg& sg = k.seg[s];
    const cuuint64_t C = sg.C, W = k.Win, H = k.Hin, N = k.N;
    CUresult r;
    if (p.mode_patch) {
      const cuuint64_t st = k.ay.it;
      cuuint64_t dims[5] = {C, st, W / st, st, N * (H / st)};
      cuuint64_t strides[4] = {C * 4, st * C * 4, W * C * 4, st * W * C * 4};
      cuuint32_t box[5] = {(cuuint32_t)TC_KB, 1, (cuuint32_t)p.TW, 1, (cuuint32_t)p.TH};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      r = enc(&maps->a[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)sg.ptr, dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[4] = {C, W, H, N};
      cuuint64_t strides[3] = {C * 4, W * C * 4, H * W * C * 4};
      cuuint32_t box[4] = {(cuuint32_t)TC_KB, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      r = enc(&maps->a[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)sg.ptr, dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return set_err(NLT_ERR_CUDA, "cuTensorMapEncodeTiled(A seg %d) failed: %d", s, (int)r);
  }
  const float* planes[2] = {bhi, blo};
  CUtensorMap* bm[2] = {&maps->bhi, &maps->blo};
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[3] = {(cuuint64_t)TC_KB, (cuuint64_t)pl.cout_pad, (cuuint64_t)p.kb_total};
    cuuint64_t strides[2] = {(cuuint64_t)TC_KB * 4, (cuuint64_t)TC_KB * 4 * pl.cout_pad};
    cuuint32_t box[3] = {(cuuint32_t)TC_KB, (cuuint32_t)pl.bn, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(bm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)planes[i], dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(NLT_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  }
  return NLT_OK;
}

// =============================================================================================
// tcgen05 weight gradient:  dW[Kd x N] = sum_p A[p, Kd]^T * G[p, N]   (3xTF32, fp32 accumulate in TMEM)
//
// The contraction runs over PIXELS.  tf32 tensor-core operands must be K-major here (the only MN-major
// tf32 layout, 128B_ATOM_32B, needs 32-channel rows and would exclude the 16-channel layers that own most
// of the bytes), so the four transform warps TRANSPOSE while they split: TMA brings raw [32 pixels x w
// channels] pieces (w <= 32, hardware-swizzled so the transposing reads are bank-conflict free), the
// transform writes hi/lo planes [channel row][32 pixels = 128 B] in the K-major SWIZZLE_128B layout, and
// the MMA warp issues 4 K-steps x 3 products per 32-pixel stage into ONE [128 x BN] accumulator that lives
// in TMEM for the CTA's whole pixel range.  Works for any channel counts that are multiples of 4.
// The bias gradient (column sums of G) is accumulated by the transform warps on the way.  One fp32
// partial per CTA goes to the workspace; the fp32 path's fixed-order reduce kernel sums the partials
// (deterministic) and scatters them into the Keras weight layout.
// =============================================================================================
constexpr int WG_PT = 32;             // pixels per stage (= 128 B of fp32 per transposed row)
constexpr int WG_MAX_PIECES = 36;     // raw pieces per stage: <= 32 for A (128 rows / 4) + <= 2 for G (+ slack)
constexpr int WG_PLANE_A = TC_BM * 128;   // one transposed A plane: 128 rows x 128 B

struct WgPiece {
  int16_t is_g, seg, uy, ux;   // source
  int16_t c0, w;               // channel range inside the source
  int16_t row, pad;            // first transposed row (A: row of the M tile; G: row of the N tile)
  int32_t raw_off;             // byte offset of the raw tile inside a raw stage (1024-aligned)
};

struct WgParams {
  int N, Hl, Wl, TW, TH, tiles_x, tiles_y, total_ptiles;
  int mode_patch, ux_step, x_off, uy_step, y_off, Hs;     // A coordinate map (as TcParams)
  int g_patch, g_px, g_py, g_Hs;                          // G on a strided sub-lattice (transposed phases); g_patch == 2: depth-to-space
  int g_ct, g_s;                                          // depth-to-space: channels of the gradient tensor, stride
  int ntap_x, nseg, ctot, Kd;
  int seg_C[NLT_MAX_SEG], seg_coff[NLT_MAX_SEG];
  int n_mtiles, n_ntiles;
  FastDiv div_tpi, div_tx;                                // multiply-shift division by tiles per image / per row
  int raw_stages, raw_stage_bytes;
  int Cout, ld, KD_pad, bias_row;
  float* ws;
};

struct WgMaps {
  CUtensorMap a[NLT_MAX_SEG][4];   // per source, per piece width class: 4, 8, 16, 32 channels
  CUtensorMap g;                   // G pieces (width BN <= 32 -> BN, else 32)
};

__device__ __forceinline__ int wg_wclass(int w) { return w == 32 ? 3 : w == 16 ? 2 : w == 8 ? 1 : 0; }

// physical 16-byte chunk of logical chunk `cq` in pixel row `p` of a raw [32 x w] tile (TMA swizzle by row bytes)
__device__ __forceinline__ int wg_raw_chunk(int w, int p, int cq) {
  return w == 32 ? (cq ^ (p & 7)) : w == 16 ? (cq ^ ((p >> 1) & 3)) : w == 8 ? (cq ^ ((p >> 2) & 1)) : cq;
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ WgMaps maps, const WgParams p) {
  constexpr int PLANE_G = BN * 128;
  constexpr int PLANES_BYTES = 2 * WG_PLANE_A + 2 * PLANE_G;      // Ahi, Alo, Ghi, Glo
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : 128;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  constexpr int GQ = BN / 4;                    // channel quads of G
  constexpr int GI = (GQ + 3) / 4;              // G quads per transform thread

  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic on the __shared__ array keeps the address space (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bars[2 * TC_MAX_STAGES + 5];
  __shared__ uint32_t tmem_base_smem;
  __shared__ WgPiece pieces[WG_MAX_PIECES];
  __shared__ uint16_t items[64];            // work items of the transform: (piece << 8) | channel quad
  __shared__ int n_pieces_s, n_items_s, raw_bytes_s;
  __shared__ float bias_part[64];

  const int S = p.raw_stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto bar_rfull = [&](int s) { return bar0 + 8u * s; };
  auto bar_rempty = [&](int s) { return bar0 + 8u * (TC_MAX_STAGES + s); };
  auto bar_tready = [&](int b) { return bar0 + 8u * (2 * TC_MAX_STAGES + b); };
  auto bar_tempty = [&](int b) { return bar0 + 8u * (2 * TC_MAX_STAGES + 2 + b); };
  const uint32_t bar_accf = bar0 + 8u * (2 * TC_MAX_STAGES + 4);
  uint8_t* planes = smem;                                         // 2 x PLANES_BYTES
  uint8_t* raw = smem + 2 * PLANES_BYTES;                         // S x raw_stage_bytes
  const uint32_t planes_u32 = smem_u32(planes), raw_u32 = smem_u32(raw);

  const int mtile = blockIdx.y, ntile = blockIdx.z;
  const int k_lo = mtile * TC_BM, k_hi = min(p.Kd, k_lo + TC_BM);

  if (threadIdx.x == 0) {
    // piece table of this (M tile, N tile): A pieces in flattened (tap, concat-channel) order, then G pieces
    int np = 0, off = 0, k = k_lo;
    while (k < k_hi) {
      const int tap = k / p.ctot, c = k - tap * p.ctot;
      int s = 0;
      while (s < p.nseg - 1 && c >= p.seg_coff[s] + p.seg_C[s]) ++s;
      const int cin = c - p.seg_coff[s];
      int w = p.seg_C[s] - cin;
      if (w > 32) w = 32;
      if (w > k_hi - k) w = k_hi - k;
      w = w >= 32 ? 32 : w >= 16 ? 16 : w >= 8 ? 8 : 4;         // power-of-two piece widths (all sizes are multiples of 4)
      WgPiece& pc = pieces[np++];
      pc.is_g = 0; pc.seg = (int16_t)s; pc.uy = (int16_t)(tap / p.ntap_x); pc.ux = (int16_t)(tap % p.ntap_x);
      pc.c0 = (int16_t)cin; pc.w = (int16_t)w; pc.row = (int16_t)(k - k_lo); pc.pad = 0; pc.raw_off = off;
      off += (WG_PT * w * 4 + 1023) & ~1023;
      k += w;
    }
    const int gw = p.g_patch == 2 ? min(min(BN, 32), p.g_ct) : (BN < 32 ? BN : 32);
    for (int c = 0; c < BN; c += gw) {
      WgPiece& pc = pieces[np++];
      pc.is_g = 1; pc.seg = 0; pc.uy = 0; pc.ux = 0; pc.c0 = (int16_t)(ntile * BN + c); pc.w = (int16_t)gw;
      if (p.g_patch == 2) {
        // depth-to-space weight gradient: GEMM column n' = tap * cout + n is channel n of the gradient pixel
        // (s*y + dy, s*x + dx) above lattice pixel (y, x)
        const int ncol = ntile * BN + c, tap = ncol / p.g_ct;
        pc.c0 = (int16_t)(ncol - tap * p.g_ct);
        pc.uy = (int16_t)(tap / p.g_s); pc.ux = (int16_t)(tap % p.g_s);
      }
      pc.row = (int16_t)c; pc.pad = 0; pc.raw_off = off;
      off += (WG_PT * gw * 4 + 1023) & ~1023;
    }
    n_pieces_s = np;
    int bytes = 0, ni = 0;
    for (int i = 0; i < np; ++i) {
      bytes += WG_PT * pieces[i].w * 4;
      for (int cq = 0; cq < pieces[i].w / 4; ++cq) items[ni++] = (uint16_t)((i << 8) | cq);
    }
    n_items_s = ni;
    raw_bytes_s = bytes;
    for (int s = 0; s < S; ++s) { mbar_init(bar_rfull(s), 1); mbar_init(bar_rempty(s), 8); }      // one arrival per transform warp
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tready(b), 8); mbar_init(bar_tempty(b), 1); }
    mbar_init(bar_accf, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // transposed planes start as zero: rows beyond this tile's Kd / BN stay zero for the whole kernel
  for (int i = threadIdx.x; i < 2 * PLANES_BYTES / 16; i += TC_THREADS)
    reinterpret_cast<float4*>(planes)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int n_pieces = n_pieces_s;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  int my_tiles = 0;
  for (int t = blockIdx.x; t < p.total_ptiles; t += gridDim.x) ++my_tiles;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_ptiles; t += gridDim.x) {
        const int n = (int)fdiv((uint32_t)t, p.div_tpi);
        const int r = t - n * tiles_per_img;
        const int rq = (int)fdiv((uint32_t)r, p.div_tx);
        const int ty0 = rq * p.TH, tx0 = (r - rq * p.tiles_x) * p.TW;
        mbar_wait(bar_rempty(stage), phase ^ 1);
        mbar_expect_tx(bar_rfull(stage), (uint32_t)raw_bytes_s);
        const uint32_t base = raw_u32 + (uint32_t)stage * p.raw_stage_bytes;
        for (int i = 0; i < n_pieces; ++i) {
          const WgPiece pc = pieces[i];
          if (pc.is_g) {
            if (p.g_patch == 2)
              tma_load_5d(base + pc.raw_off, &maps.g, bar_rfull(stage), pc.c0, pc.ux, tx0, pc.uy, n * p.g_Hs + ty0);
            else if (p.g_patch)
              tma_load_5d(base + pc.raw_off, &maps.g, bar_rfull(stage), pc.c0, p.g_px, tx0, p.g_py, n * p.g_Hs + ty0);
            else
              tma_load_4d(base + pc.raw_off, &maps.g, bar_rfull(stage), pc.c0, tx0, ty0, n);
          } else {
            const CUtensorMap* m = &maps.a[pc.seg][wg_wclass(pc.w)];
            if (p.mode_patch)
              tma_load_5d(base + pc.raw_off, m, bar_rfull(stage), pc.c0, pc.ux, tx0, pc.uy, n * p.Hs + ty0);
            else
              tma_load_4d(base + pc.raw_off, m, bar_rfull(stage), pc.c0, tx0 + pc.ux * p.ux_step + p.x_off,
                          ty0 + pc.uy * p.uy_step + p.y_off, n);
          }
        }
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      for (int i = 0; i < my_tiles; ++i) {
        const int b = i & 1;
        mbar_wait(bar_tready(b), (uint32_t)(i >> 1) & 1);
        tc_fence_after();
        const uint32_t ahi = planes_u32 + (uint32_t)b * PLANES_BYTES, alo = ahi + WG_PLANE_A;
        const uint32_t ghi = alo + WG_PLANE_A, glo = ghi + PLANE_G;
#pragma unroll
        for (int ks = 0; ks < WG_PT / 8; ++ks) {
          const uint64_t dah = umma_desc_kmajor<128>(ahi + ks * 32), dal = umma_desc_kmajor<128>(alo + ks * 32);
          const uint64_t dgh = umma_desc_kmajor<128>(ghi + ks * 32), dgl = umma_desc_kmajor<128>(glo + ks * 32);
          tc_mma_tf32(tmem_base, dal, dgh, IDESC, (i | ks) != 0);
          tc_mma_tf32(tmem_base, dah, dgl, IDESC, 1);
          tc_mma_tf32(tmem_base, dah, dgh, IDESC, 1);
        }
        tc_commit(bar_tempty(b));
        if (i == my_tiles - 1) tc_commit(bar_accf);
      }
    }
  } else {
    // ================= transpose + split: all eight remaining warps (2..9) =================
    // ncu: with four warps this role ran one warp per scheduler at ~0.3 IPC (dependent LDS -> cvt -> STS
    // chains) and set the stage time; eight warps and two independent items per thread hide that latency.
    const int t = threadIdx.x - 64;               // 0..255 ; pixel = t % 32, item lane = t / 32
    const int px = t & 31, ql = t >> 5;
    const int n_items = n_items_s;
    constexpr int MAXI = 6;                       // items per thread: 48 items / 8 lanes
    float gsum[MAXI][4];
#pragma unroll
    for (int i = 0; i < MAXI; ++i) { gsum[i][0] = gsum[i][1] = gsum[i][2] = gsum[i][3] = 0.f; }
    const int pxc = px >> 2, pxo = (px & 3) << 2;
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      mbar_wait(bar_rfull(stage), phase);
      mbar_wait(bar_tempty(b), ((uint32_t)(it >> 1) & 1) ^ 1);   // MMAs of two stages ago are done with planes b
      const uint8_t* rbase = raw + (size_t)stage * p.raw_stage_bytes;
      uint8_t* pl = planes + (size_t)b * PLANES_BYTES;
#pragma unroll
      for (int i = 0; i < MAXI; ++i) {
        const int j = ql + 8 * i;
        if (j >= n_items) break;
        const int itm = items[j];
        const WgPiece pc = pieces[itm >> 8];
        const int cq = itm & 0xff;
        uint8_t* hi = pl + (pc.is_g ? 2 * WG_PLANE_A : 0);
        const int lo_off = pc.is_g ? PLANE_G : WG_PLANE_A;
        const float4 v = *reinterpret_cast<const float4*>(rbase + pc.raw_off + (size_t)px * pc.w * 4 +
                                                          wg_raw_chunk(pc.w, px, cq) * 16);
        const float vv[4] = {v.x, v.y, v.z, v.w};
        if (pc.is_g) { gsum[i][0] += v.x; gsum[i][1] += v.y; gsum[i][2] += v.z; gsum[i][3] += v.w; }
        const int row0 = pc.row + cq * 4;           // multiple of 4: (row0 + e) & 7 == (row0 & 7) + e
        const int r7 = row0 & 7;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // K-major SWIZZLE_128B: 16-byte chunk index XOR (row % 8)
          const int o = (row0 + e) * 128 + (((pxc ^ (r7 + e)) << 4) | pxo);
          const float h = tf32_round(vv[e]);
          *reinterpret_cast<float*>(hi + o) = h;
          *reinterpret_cast<float*>(hi + lo_off + o) = tf32_round(vv[e] - h);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_tready(b));
        mbar_arrive(bar_rempty(stage));
      }
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
    // bias partial: reduce the 32 pixel lanes of each warp in a fixed butterfly order
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int j = ql + 8 * i;
      const bool isg = j < n_items && pieces[items[j < n_items ? j : 0] >> 8].is_g;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = gsum[i][e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (isg && px == 0) {
          const int itm = items[j];
          bias_part[pieces[itm >> 8].row + (itm & 0xff) * 4 + e] = v;
        }
      }
    }
  }
  // ---- all partials are ready: accumulator (after accf) and bias sums ----
  __syncthreads();
  if (warp >= 6) {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;            // accumulator row
    const int k = k_lo + r;
    if (my_tiles > 0) {
      mbar_wait(bar_accf, 0);
      tc_fence_after();
    }
    float* dst = p.ws + ((size_t)blockIdx.x * p.KD_pad + (k < k_hi ? k : 0)) * p.ld + ntile * BN;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      if (my_tiles > 0) {
        tc_ld16(taddr + c0, v);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = 0u;
      }
      if (k < k_hi) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          *reinterpret_cast<float4*>(dst + c0 + q4 * 4) =
              make_float4(__uint_as_float(v[q4 * 4]), __uint_as_float(v[q4 * 4 + 1]), __uint_as_float(v[q4 * 4 + 2]),
                          __uint_as_float(v[q4 * 4 + 3]));
      }
    }
    if (mtile == 0 && r < BN)
      p.ws[((size_t)blockIdx.x * p.KD_pad + p.bias_row) * p.ld + ntile * BN + r] = my_tiles > 0 ? bias_part[r] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

struct WgPlan {
  bool ok;
  int bn;
  WgParams p;
  size_t smem_bytes;
  int nsplit;
};

static WgPlan wg_plan(const GConvK& k) {
  WgPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (k.M == 0 || k.Cout % 16 != 0) return pl;
  if (k.d2s && (k.cout_true % 4 != 0 || (k.cout_true & (k.cout_true - 1)) != 0) && k.cout_true % 32 != 0) return pl;
  int ctot = 0;
  for (int s = 0; s < k.nseg; ++s) {
    const Seg& sg = k.seg[s];
    if (sg.C % 4 != 0 || sg.sub != nullptr || sg.bcast || !aligned16(sg.ptr)) return pl;
    ctot += sg.C;
  }
  if (k.ay.nu < 1 || k.ax.nu < 1) return pl;
  pl.bn = k.Cout % 64 == 0 ? 64 : k.Cout % 32 == 0 ? 32 : 16;
  WgParams& p = pl.p;
  p.N = k.N; p.Hl = k.ay.nt; p.Wl = k.ax.nt;
  if (p.Wl >= WG_PT) {
    if (p.Wl % WG_PT) return pl;
    p.TW = WG_PT; p.TH = 1;
  } else {
    if (WG_PT % p.Wl) return pl;
    p.TW = p.Wl; p.TH = WG_PT / p.Wl;
    if (p.Hl % p.TH) return pl;
  }
  p.tiles_x = p.Wl / p.TW; p.tiles_y = p.Hl / p.TH;
  p.div_tpi = make_fastdiv((uint32_t)(p.tiles_x * p.tiles_y)); p.div_tx = make_fastdiv((uint32_t)p.tiles_x);
  const long long tt = (long long)p.N * p.tiles_x * p.tiles_y;
  if (tt > (1ll << 30) || tt < 4) return pl;
  p.total_ptiles = (int)tt;
  const bool stride1 = (k.ay.it == 1 && k.ax.it == 1);
  const bool patch = (k.ay.it > 1 && k.ay.it == k.ax.it && k.ay.iu == 1 && k.ax.iu == 1 && k.ay.i0 == 0 && k.ax.i0 == 0 &&
                      k.ay.nu == k.ay.it && k.ax.nu == k.ax.it && k.Hin == k.ay.nt * k.ay.it && k.Win == k.ax.nt * k.ax.it);
  if (!stride1 && !patch) return pl;
  p.mode_patch = patch ? 1 : 0;
  p.ux_step = k.ax.iu; p.x_off = k.ax.i0; p.uy_step = k.ay.iu; p.y_off = k.ay.i0;
  p.Hs = patch ? k.Hin / k.ay.it : 0;
  if (k.ay.os != k.ax.os) return pl;
  if (k.d2s) {
    // ONE pass over the input lattice for all s*s taps (instead of s*s phases that each re-read the input):
    // the gradient pieces of a stage come from the s*s pixels above each lattice pixel
    if (k.Hout != p.Hl * k.d2s_s || k.Wout != p.Wl * k.d2s_s) return pl;
    p.g_patch = 2; p.g_py = 0; p.g_px = 0; p.g_Hs = k.Hout / k.d2s_s; p.g_ct = k.cout_true; p.g_s = k.d2s_s;
  } else if (k.ay.os == 1) {
    if (k.ay.o0 != 0 || k.ax.o0 != 0 || k.Hout != p.Hl || k.Wout != p.Wl) return pl;
    p.g_patch = 0;
  } else {
    const int os = k.ay.os;
    if (k.Hout != p.Hl * os || k.Wout != p.Wl * os) return pl;
    p.g_patch = 1; p.g_py = k.ay.o0; p.g_px = k.ax.o0; p.g_Hs = k.Hout / os;
  }
  p.ntap_x = k.ax.nu; p.nseg = k.nseg; p.ctot = ctot;
  p.Kd = k.ay.nu * k.ax.nu * ctot;
  if (p.Kd < 128) return pl;      // measured (also with 8 transform warps): below one full M tile the warp-stream fp32 kernel wins
  int coff = 0;
  for (int s = 0; s < k.nseg; ++s) { p.seg_C[s] = k.seg[s].C; p.seg_coff[s] = coff; coff += k.seg[s].C; }
  p.n_mtiles = (p.Kd + TC_BM - 1) / TC_BM;
  p.n_ntiles = k.Cout / pl.bn;
  p.Cout = k.Cout; p.ld = k.Cout;
  p.bias_row = p.Kd; p.KD_pad = p.Kd + 4;
  // raw stage: worst case 32 A pieces of 4 channels (1 KB each after rounding) or 4 of 32 (4 KB each): <= 32 KB; G <= 8 KB
  p.raw_stage_bytes = 32 * 1024 + 8 * 1024;
  {
    // tighter bound: pieces never exceed max(16 KB of payload, #pieces KB); recompute exactly for M tile 0 (the fullest)
    int bytes = 0, kk = 0;
    const int k_hi = p.Kd < TC_BM ? p.Kd : TC_BM;
    int worst = 0;
    for (int mt = 0; mt < p.n_mtiles; ++mt) {
      bytes = 0; kk = mt * TC_BM;
      const int hi = (kk + TC_BM < p.Kd) ? kk + TC_BM : p.Kd;
      while (kk < hi) {
        const int tap = kk / ctot, c = kk - tap * ctot;
        int s = 0;
        while (s < k.nseg - 1 && c >= p.seg_coff[s] + p.seg_C[s]) ++s;
        int w = p.seg_C[s] - (c - p.seg_coff[s]);
        if (w > 32) w = 32;
        if (w > hi - kk) w = hi - kk;
        w = w >= 32 ? 32 : w >= 16 ? 16 : w >= 8 ? 8 : 4;
        bytes += (WG_PT * w * 4 + 1023) & ~1023;
        kk += w;
      }
      if (bytes > worst) worst = bytes;
    }
    (void)k_hi;
    int gw = pl.bn < 32 ? pl.bn : 32;
    if (k.d2s && k.cout_true < gw) gw = k.cout_true;
    worst += (pl.bn / gw) * ((WG_PT * gw * 4 + 1023) & ~1023);
    p.raw_stage_bytes = worst;
  }
  const size_t planes = 2 * (2 * (size_t)WG_PLANE_A + 2 * (size_t)pl.bn * 128);
  int st = (int)((227 * 1024 - 2048 - (long long)planes - 4096) / p.raw_stage_bytes);
  if (st > TC_MAX_STAGES) st = TC_MAX_STAGES;
  if (st < 2) return pl;
  p.raw_stages = st;
  pl.smem_bytes = planes + (size_t)st * p.raw_stage_bytes + 1024;
  long long want = 148 / ((long long)p.n_mtiles * p.n_ntiles);
  if (want < 1) want = 1;
  if (want > tt / 4) want = tt / 4;
  if (want < 1) want = 1;
  pl.nsplit = (int)want;
  if (get_encode() == nullptr) return pl;
  pl.ok = true;
  return pl;
}

bool tc_wgrad_applicable(const GConvK& k) { return wg_plan(k).ok; }

size_t tc_wgrad_ws_floats(const GConvK& k) {
  WgPlan pl = wg_plan(k);
  return pl.ok ? (size_t)pl.nsplit * pl.p.KD_pad * pl.p.ld : 0;
}

template <int BN>
static int wg_launch(const WgMaps& maps, const WgPlan& pl, cudaStream_t st) {
  static bool attr_set[64] = {false};     // per device: the attribute is not process-wide
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096);
    if (e != cudaSuccess) return set_err(NLT_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  dim3 grid(pl.nsplit, pl.p.n_mtiles, pl.p.n_ntiles);
  tc_wgrad_kernel<BN><<<grid, TC_THREADS, pl.smem_bytes, st>>>(maps, pl.p);
  NLT_CUDA_LAUNCH_CHECK("tc_wgrad_kernel");
  __atomic_add_fetch(&g_tc_launches, 1ull, __ATOMIC_RELAXED);
  return NLT_OK;
}

// raw [pixels x w channels] tile of an NHWC tensor, swizzled by its row bytes (bank-conflict-free transposing reads)
static CUresult wg_encode(EncodeTiledFn enc, CUtensorMap* m, const float* ptr, int C, int W, int H, int N, int w, int TW,
                          int TH, int patch_s) {
  const CUtensorMapSwizzle swz = w == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : w == 16 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : w == 8 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  const cuuint64_t c = C, ww = W, h = H, n = N;
  if (patch_s > 1) {
    const cuuint64_t st_ = patch_s;
    cuuint64_t dims[5] = {c, st_, ww / st_, st_, n * (h / st_)};
    cuuint64_t strides[4] = {c * 4, st_ * c * 4, ww * c * 4, st_ * ww * c * 4};
    cuuint32_t box[5] = {(cuuint32_t)w, 1, (cuuint32_t)TW, 1, (cuuint32_t)TH};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  cuuint64_t dims[4] = {c, ww, h, n};
  cuuint64_t strides[3] = {c * 4, ww * c * 4, h * ww * c * 4};
  cuuint32_t box[4] = {(cuuint32_t)w, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int launch_tc_wgrad(const GConvK& k, const float* G, float* ws, WgradK* w, size_t* KD_pad, cudaStream_t st) {
  WgPlan pl = wg_plan(k);
  if (!pl.ok) return set_err(NLT_ERR_INVALID, "tensor-core wgrad not applicable");
  NLT_CHECK_ARG(aligned16(G), "dz must be 16-byte aligned");
  EncodeTiledFn enc = get_encode();
  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  const WgParams& p = pl.p;
  static const int widths[4] = {4, 8, 16, 32};
  for (int s = 0; s < k.nseg; ++s)
    for (int wi = 0; wi < 4; ++wi) {
      if (widths[wi] > k.seg[s].C) continue;
      CUresult r = wg_encode(enc, &maps.a[s][wi], k.seg[s].ptr, k.seg[s].C, k.Win, k.Hin, k.N, widths[wi], p.TW, p.TH,
                             p.mode_patch ? k.ay.it : 1);
      if (r != CUDA_SUCCESS) return set_err(NLT_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad A %d/%d) failed: %d", s, widths[wi], (int)r);
    }
  {
    int gw = pl.bn < 32 ? pl.bn : 32;
    if (k.d2s && k.cout_true < gw) gw = k.cout_true;
    CUresult r = k.d2s ? wg_encode(enc, &maps.g, G, k.cout_true, k.Wout, k.Hout, k.N, gw, p.TW, p.TH, k.d2s_s)
                       : wg_encode(enc, &maps.g, G, k.Cout, k.Wout, k.Hout, k.N, gw, p.TW, p.TH, p.g_patch ? k.ay.os : 1);
    if (r != CUDA_SUCCESS) return set_err(NLT_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad G) failed: %d", (int)r);
  }
  pl.p.ws = ws;
  w->g = k;
  w->GS = p.ctot / 4;
  w->KG = k.ay.nu * k.ax.nu * w->GS + 1;
  w->ld = p.ld; w->nsplit = pl.nsplit; w->pix_per_split = 0;
  *KD_pad = (size_t)p.KD_pad;
  switch (pl.bn) {
    case 64: return wg_launch<64>(maps, pl, st);
    case 32: return wg_launch<32>(maps, pl, st);
    default: return wg_launch<16>(maps, pl, st);
  }
}

bool tc_applicable(const GConvK& k) { return tc_plan(k).ok; }

size_t tc_workspace_bytes(const GConvK& k) {
  TcPlan pl = tc_plan(k);
  return pl.ok ? 2 * pl.pack_floats * sizeof(float) + 256 : 0;
}

template <int BN, int KBW>
static int tc_launch_bn(const TcMaps& maps, const TcPlan& pl, cudaStream_t st) {
  // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute
  static bool attr_set[2][64] = {{false}};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[pl.ts ? 1 : 0][dev]) {
    cudaError_t e = pl.ts ? cudaFuncSetAttribute(tcs_gconv_kernel<BN, KBW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 227 * 1024 - 1024)
                          : cudaFuncSetAttribute(tc_gconv_kernel<BN, KBW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 227 * 1024 - 1024);
    if (e != cudaSuccess) return set_err(NLT_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[pl.ts ? 1 : 0][dev] = true;
  }
  const int grid = pl.p.total_tiles < 148 ? pl.p.total_tiles : 148;
  if (pl.ts) tcs_gconv_kernel<BN, KBW><<<grid, TCF_THREADS, pl.smem_bytes, st>>>(maps, pl.p);
  else tc_gconv_kernel<BN, KBW><<<grid, TCF_THREADS, pl.smem_bytes, st>>>(maps, pl.p);
  NLT_CUDA_LAUNCH_CHECK("tc_gconv_kernel");
  __atomic_add_fetch(&g_tc_launches, 1ull, __ATOMIC_RELAXED);
  return NLT_OK;
}

// hi / lo planes of the weights in K-block layout (what the B tensor maps of the conv kernel read)
int tc_pack(const GConvK& k, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  TcPlan pl = tc_plan(k);
  if (!pl.ok) return set_err(NLT_ERR_INVALID, "tensor-core path not applicable");
  const size_t need = 2 * pl.pack_floats * sizeof(float) + 256;
  NLT_CHECK_ARG(workspace != nullptr && workspace_bytes >= need, "tc workspace too small: need %zu", need);
  float* bhi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  float* blo = bhi + pl.pack_floats;
  const size_t total = pl.pack_floats;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  tc_pack_weights_kernel<<<blocks, 256, 0, st>>>(k, pl.kbw, pl.p.kb_total, pl.chunks_per_tap, pl.cout_pad, bhi, blo);
  NLT_CUDA_LAUNCH_CHECK("tc_pack_weights_kernel");
  return NLT_OK;
}

int launch_tc(const GConvK& k, const float* bias, int act, float beta, const float* mask_y, int mask_act, float* out,
              void* workspace, size_t workspace_bytes, cudaStream_t st, bool prepacked) {
  TcPlan pl = tc_plan(k);
  if (!pl.ok) return set_err(NLT_ERR_INVALID, "tensor-core path not applicable");
  const size_t need = 2 * pl.pack_floats * sizeof(float) + 256;
  NLT_CHECK_ARG(workspace != nullptr && workspace_bytes >= need, "tc workspace too small: need %zu", need);
  float* bhi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  float* blo = bhi + pl.pack_floats;
  if (!prepacked) {
    const int rc0 = tc_pack(k, workspace, workspace_bytes, st);
    if (rc0 != NLT_OK) return rc0;
  }
  TcMaps maps;
  int rc = encode_maps(k, pl, bhi, blo, &maps);
  if (rc != NLT_OK) return rc;
  pl.p.act = act; pl.p.mask_act = mask_act; pl.p.beta = beta; pl.p.bias = bias; pl.p.mask_y = mask_y; pl.p.out = out;
  if (pl.kbw == 32) {
    switch (pl.bn) {
      case 128: return tc_launch_bn<128, 32>(maps, pl, st);
      case 64: return tc_launch_bn<64, 32>(maps, pl, st);
      case 32: return tc_launch_bn<32, 32>(maps, pl, st);
      default: return tc_launch_bn<16, 32>(maps, pl, st);
    }
  }
  switch (pl.bn) {
    case 128: return tc_launch_bn<128, 16>(maps, pl, st);
    case 64: return tc_launch_bn<64, 16>(maps, pl, st);
    case 32: return tc_launch_bn<32, 16>(maps, pl, st);
    default: return tc_launch_bn<16, 16>(maps, pl, st);
  }
}

}  // namespace nlt
