// bp_gemm.cuh -- tcgen05 GEMM family over "batch-planar" tensors for sm_100a.
//
// Layout.  Every tensor on the tensor-core path (decoded frames, activations, activation gradients, weights) is
// stored as two bf16 planes hi = bf16(x), lo = bf16(x - hi) ("bf16x3": hi*hi + hi*lo + lo*hi in fp32 TMEM
// accumulators reproduces fp32 products to ~2^-17) in the batch-planar order
//
//        element (row r, feature f)  ->  plane[ ((f >> 3) * pitch + r) * 8 + (f & 7) ]
//
// i.e. [feature chunk of 8][row][8]: for activations the row is the SAMPLE and a feature is (pixel, channel) in HWC
// order; for a weight matrix W[K, N] the row is k and the feature is n.  Consequences:
//   * the 16-byte pieces of consecutive rows are adjacent, so ANY operand tile -- the 128 samples x 64 channels of
//     one filter tap, the 64 samples x 128 features of a weight-gradient operand, a weight tile -- is a handful of
//     contiguous runs that the TMA engine moves with cp.async.bulk (no gather, no im2col copy, no tensor map), and
//     lands in shared memory directly in the UMMA no-swizzle canonical layout (8 rows x 16 B core matrices);
//   * an epilogue thread owns one sample row, and the 32 lanes of a warp store 512 contiguous bytes per chunk;
//   * a convolution is a GEMM per output pixel whose M rows are the samples and whose K loop walks the filter
//     taps: stride, padding (out-of-image taps are simply skipped) and the transposed convolution of the data
//     gradient are all "which feature chunks does this stage read" -- no zero-filled operand tiles, no parity classes.
//
// Kernels (all persistent, warp specialised: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 =
// epilogue, double-buffered TMEM accumulators, mbarrier full/empty rings):
//   bp_rows_kernel<0>  forward     D[b, n]  = act(alpha * sum_k A[b, k] W[k, n] + bias)        conv / dense
//   bp_rows_kernel<1>  forward, split-K partial sums (dense layers with a long K)
//   bp_rows_kernel<2>  data grad   D[b, k]  = (sum_n G[b, n] W[k, n]) * act'(X[b, k])           conv / dense
//   bp_wgrad_kernel    weight grad D[k, n]  = sum_b X[b, k] G[b, n]   (samples are the reduction axis; conv layers
//                      accumulate one TMEM accumulator per filter row over all output pixels of the CTA and write
//                      per-CTA partial sums that grad_reduce_kernel adds in a fixed order: no atomics)
// UMMA descriptors (SWIZZLE_NONE):  K-major operand  LBO = chunk-plane stride, SBO = 128 B, K step = 2 planes;
//                                   MN-major operand LBO = 128 B, SBO = chunk-plane stride, K step = 256 B.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_f32.cuh"

namespace xtb {
namespace bp {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(20000u)
      : "memory");
  return ok;
}
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 20)) __trap();
  }
}
// TMA bulk copy global -> shared (1-D, no tensor map); completion is counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns; the caller waits with tmem_wait_ld() before touching v
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, SWIZZLE_NONE (sm_100 descriptor version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: bf16 x bf16 -> fp32, M = 128, N = n; major bits: 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// hi = bf16x2(a,b); lo = bf16x2(a - float(hi.a), b - float(hi.b)); a bf16 widened to fp32 is its bits << 16
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);
  float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
  lo = pack_bf16(a - ha, b - hb);
}
__device__ __forceinline__ void unpack8(const uint4& u, float v[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void split8(const float v[8], uint4& hi, uint4& lo) {
  split2(v[0], v[1], hi.x, lo.x);
  split2(v[2], v[3], hi.y, lo.y);
  split2(v[4], v[5], hi.z, lo.z);
  split2(v[6], v[7], hi.w, lo.w);
}

// ------------------------------------------------------------------------------------------
// batch-planar tensor handle
// ------------------------------------------------------------------------------------------
struct BpT {
  bf16* hi;            // hi plane; NULL = absent
  long long lo_off;    // lo plane = hi + lo_off (elements)
  int pitch;           // rows per feature chunk (multiple of 16)
};
__host__ __device__ inline long long bp_index(int pitch, int row, int f) { return ((long long)(f >> 3) * pitch + row) * 8 + (f & 7); }

// ------------------------------------------------------------------------------------------
// bp_rows_kernel: the M rows of a tile are 128 samples; a "unit" is what one accumulator tile produces
//   conv forward  : unit = output pixel, N = Cout           K stages walk the filter rows of that pixel
//   conv data grad: unit = input pixel,  N = Cin            K stages walk the taps that reach that pixel
//   dense         : unit = (N tile, K split)                K stages walk a range of feature chunks
// Weights ("blob"): batch-planar W^T, i.e. [n chunk][k row][8] with pitch = K rows.  Forward reads it MN-major (rows k,
// chunks along n), the data gradient K-major (rows = output feature k, chunks along the reduction n): same bytes.
// ------------------------------------------------------------------------------------------
constexpr int RW_THREADS = 192;
constexpr int RW_A_PLANE = 2048;              // 128 rows x 16 B
constexpr int RW_STAGE_A = 2 * 8 * RW_A_PLANE;   // hi + lo, 8 chunks (64 K elements)
constexpr int RW_STAGE_B = 16384;
constexpr int RW_MAX_STAGES = 4;

struct RowsArgs {
  BpT a; int a_split;                               // operand rows (activations / gradients); a_split: lo plane is read
  const bf16* w_hi; const bf16* w_lo; int w_pitch;  // weight blob planes
  int w_res; int w_res_chunks;                      // blob resident in shared memory (conv): chunks to load per plane
  int mode;                                         // 0 conv forward, 1 conv data gradient, 2 dense
  int H, W, C, OH, OW, Cout, KH, KW, S, padT, padL; // conv geometry: [H,W,C] -> [OH,OW,Cout]
  int kchunks, kc_split, n_ntiles;                  // dense: K chunks in total / per split, N tiles
  int n_units, n_btiles, B, N;                      // tiles = n_units * n_btiles; N = accumulator columns per plane
  // forward epilogue
  BpT out; float* out_f32; int ld_f32; const float* bias; float alpha; int act;
  float* part; long long part_z; int ld_part;       // split-K partial sums part[z][b][ld_part]
  // data-gradient epilogue: out = acc * act'(src) (+ out); column sums into db_part[cta][N] when non-NULL
  BpT src; int src_act; int accumulate; float* db_part;
};

struct StageDesc { int a_chunk0, nch, w_row, w_c0; };
struct ItState { int y, x, ky, t, chunk, rem, wrow, end, ntile; };

template <int KIND>
__device__ __forceinline__ void it_init(const RowsArgs& a, int u, ItState& s) {
  if (a.mode == 2) {
    s.ntile = u % a.n_ntiles;
    int z = u / a.n_ntiles;
    s.chunk = z * a.kc_split;
    s.end = min(a.kchunks, s.chunk + a.kc_split);
  } else {
    int wu = (KIND == 2) ? a.W : a.OW;
    s.y = u / wu; s.x = u - s.y * wu;
    s.ky = -1; s.rem = 0; s.t = -1;
  }
}
template <int KIND>
__device__ __forceinline__ bool it_next(const RowsArgs& a, ItState& s, StageDesc& d) {
  if (a.mode == 2) {
    if (s.chunk >= s.end) return false;
    d.a_chunk0 = s.chunk; d.nch = min(8, s.end - s.chunk);
    if (KIND == 2) { d.w_row = s.ntile * a.N; d.w_c0 = s.chunk; }
    else { d.w_row = s.chunk * 8; d.w_c0 = s.ntile * (a.N >> 3); }
    s.chunk += d.nch;
    return true;
  }
  if (KIND != 2) {
    // forward conv: per filter row the taps inside the image are one contiguous feature run
    if (s.rem == 0) {
      for (;;) {
        if (++s.ky >= a.KH) return false;
        int iy = s.y * a.S - a.padT + s.ky;
        if (iy < 0 || iy >= a.H) continue;
        int x0 = s.x * a.S - a.padL, xl = max(x0, 0), xh = min(x0 + a.KW, a.W);
        if (xh <= xl) continue;
        s.chunk = ((iy * a.W + xl) * a.C) >> 3;
        s.rem = ((xh - xl) * a.C) >> 3;
        s.wrow = (s.ky * a.KW + (xl - x0)) * a.C;
        break;
      }
    }
    d.a_chunk0 = s.chunk; d.nch = min(8, s.rem); d.w_row = s.wrow; d.w_c0 = 0;
    s.chunk += d.nch; s.rem -= d.nch; s.wrow += d.nch * 8;
    return true;
  }
  // data gradient: one stage per filter tap whose output position exists
  for (;;) {
    if (++s.t >= a.KH * a.KW) return false;
    int ky = s.t / a.KW, kx = s.t - ky * a.KW;
    int ty = s.y + a.padT - ky, tx = s.x + a.padL - kx;
    if (ty < 0 || tx < 0) continue;
    int oy = ty / a.S, ox = tx / a.S;
    if (oy * a.S != ty || ox * a.S != tx || oy >= a.OH || ox >= a.OW) continue;
    d.a_chunk0 = ((oy * a.OW + ox) * a.Cout) >> 3;
    d.nch = a.Cout >> 3;
    d.w_row = s.t * a.C;
    d.w_c0 = 0;
    return true;
  }
}

template <int KIND>
__global__ void __launch_bounds__(RW_THREADS, 1)
bp_rows_kernel(const __grid_constant__ RowsArgs a, int n_stages, int stage_bytes, int wres_bytes, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ __align__(8) uint64_t bars[2 * RW_MAX_STAGES + 5];
  __shared__ uint32_t tmem_slot;
  __shared__ float red_sh[4][64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[RW_MAX_STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * RW_MAX_STAGES]), tempty0 = smem_u32(&bars[2 * RW_MAX_STAGES + 2]);
  const uint32_t wbar = smem_u32(&bars[2 * RW_MAX_STAGES + 4]);
  if (tid == 0) {
    for (int s = 0; s < RW_MAX_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; b++) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 4); }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t wres = smem_u32(smem);                  // resident weights: hi plane chunks then lo plane chunks
  const uint32_t stage0 = wres + (uint32_t)wres_bytes;
  const int total = a.n_units * a.n_btiles;
  const int N = a.N;
  constexpr bool CAT = (KIND != 2);                      // forward: one MMA over [W_hi | W_lo], accumulator 2N columns
  const int acc_cols = CAT ? 2 * N : N;

  if (warp == 0) {
    // ================= TMA producer =================
    if (a.w_res && lane == 0) {
      const uint32_t bytes = (uint32_t)a.w_res_chunks * a.w_pitch * 16;
      mbar_expect_tx(wbar, 2 * bytes);
      bulk_g2s(wres, a.w_hi, bytes, wbar);
      bulk_g2s(wres + bytes, a.w_lo, bytes, wbar);
    }
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int u = tile / a.n_btiles, bt = tile - u * a.n_btiles;
      const int b0 = bt * 128, nr = min(128, a.B - b0);
      ItState it; StageDesc d;
      it_init<KIND>(a, u, it);
      while (it_next<KIND>(a, it, d)) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const int n_a = d.nch * (a.a_split ? 2 : 1);
        const int n_bh = a.w_res ? 0 : (KIND == 2 ? d.nch : (N >> 3));
        const uint32_t a_bytes = (uint32_t)nr * 16;
        const uint32_t b_bytes = (KIND == 2) ? (uint32_t)N * 16 : (uint32_t)d.nch * 128;
        if (lane == 0) mbar_expect_tx(full0 + 8 * stage, a_bytes * n_a + b_bytes * 2 * n_bh);
        __syncwarp();
        const uint32_t sA = stage0 + (uint32_t)stage * stage_bytes, sB = sA + RW_STAGE_A;
        for (int i = lane; i < n_a + 2 * n_bh; i += 32) {
          if (i < n_a) {
            const int pl = i >= d.nch, c = i - pl * d.nch;
            const bf16* src = a.a.hi + (pl ? a.a.lo_off : 0) + ((long long)(d.a_chunk0 + c) * a.a.pitch + b0) * 8;
            bulk_g2s(sA + pl * (8 * RW_A_PLANE) + c * RW_A_PLANE, src, a_bytes, full0 + 8 * stage);
          } else {
            const int j = i - n_a, pl = j >= n_bh, c = j - pl * n_bh;
            const bf16* src = (pl ? a.w_lo : a.w_hi) + ((long long)(d.w_c0 + c) * a.w_pitch + d.w_row) * 8;
            const uint32_t dst = (KIND == 2) ? sB + pl * 8192 + c * (N * 16) : sB + (pl * (N >> 3) + c) * 1024;
            bulk_g2s(dst, src, b_bytes, full0 + 8 * stage);
          }
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc_n = make_idesc(N, false, !(KIND == 2));
    const uint32_t idesc_cat = make_idesc(2 * N, false, true);
    if (a.w_res) mbar_wait(wbar, 0);
    int stage = 0; uint32_t phase = 0; int cnt = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, cnt++) {
      const int u = tile / a.n_btiles;
      const int buf = cnt & 1;
      const uint32_t acc = tmem_base + buf * acc_cols;
      mbar_wait(tempty0 + 8 * buf, ((cnt >> 1) & 1) ^ 1);
      tc_fence_after();
      ItState it; StageDesc d;
      it_init<KIND>(a, u, it);
      uint32_t accum = 0;
      while (it_next<KIND>(a, it, d)) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = stage0 + (uint32_t)stage * stage_bytes, sB = sA + RW_STAGE_A;
          const uint32_t sA_lo = sA + 8 * RW_A_PLANE;
          for (int j = 0; j < (d.nch >> 1); j++) {
            const uint64_t ah = make_desc(sA + j * 2 * RW_A_PLANE, RW_A_PLANE, 128);
            if (KIND != 2) {
              // W MN-major: rows k (16 per step), chunks along n; [hi | lo] chunk planes are adjacent
              const uint64_t bw = a.w_res ? make_desc(wres + (uint32_t)(d.w_row + 16 * j) * 16, 128, (uint32_t)a.w_pitch * 16)
                                          : make_desc(sB + j * 256, 128, 1024);
              umma_bf16(acc, ah, bw, idesc_cat, accum);
              if (a.a_split) umma_bf16(acc, make_desc(sA_lo + j * 2 * RW_A_PLANE, RW_A_PLANE, 128), bw, idesc_n, 1u);
            } else {
              // W K-major: rows = output feature, chunks along the reduction (2 per step)
              uint64_t bh, bl;
              if (a.w_res) {
                const uint32_t off = ((uint32_t)(d.w_c0 + 2 * j) * a.w_pitch + d.w_row) * 16;
                const uint32_t plane = (uint32_t)a.w_res_chunks * a.w_pitch * 16;
                bh = make_desc(wres + off, (uint32_t)a.w_pitch * 16, 128);
                bl = make_desc(wres + plane + off, (uint32_t)a.w_pitch * 16, 128);
              } else {
                bh = make_desc(sB + j * 2 * (N * 16), (uint32_t)N * 16, 128);
                bl = make_desc(sB + 8192 + j * 2 * (N * 16), (uint32_t)N * 16, 128);
              }
              umma_bf16(acc, ah, bh, idesc_n, accum);
              umma_bf16(acc, ah, bl, idesc_n, 1u);
              if (a.a_split) umma_bf16(acc, make_desc(sA_lo + j * 2 * RW_A_PLANE, RW_A_PLANE, 128), bh, idesc_n, 1u);
            }
            accum = 1u;
          }
          umma_commit(empty0 + 8 * stage);
        }
        __syncwarp();
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) umma_commit(tfull0 + 8 * buf);
      __syncwarp();
    }
  } else {
    // ================= epilogue (warps 2..5; warp & 3 = TMEM lane quarter) =================
    const int q = warp & 3;
    float dbacc[64];
    if (KIND == 2) {
#pragma unroll
      for (int j = 0; j < 64; j++) dbacc[j] = 0.f;
    }
    const int b_pad = (a.B + 15) & ~15;
    int cnt = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, cnt++) {
      const int u = tile / a.n_btiles, bt = tile - u * a.n_btiles;
      const int buf = cnt & 1;
      const int b = bt * 128 + q * 32 + lane;
      int oc0, z = 0;                          // first output chunk of the unit
      if (a.mode == 2) { const int nt = u % a.n_ntiles; z = u / a.n_ntiles; oc0 = nt * (N >> 3); }
      else oc0 = u * (N >> 3);
      mbar_wait(tfull0 + 8 * buf, (cnt >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * acc_cols;
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        if (c < N) {
          uint32_t r0[16], r1[16];
          tmem_ld16_nowait(taddr + c, r0);
          if (CAT) tmem_ld16_nowait(taddr + N + c, r1);
          tmem_wait_ld();
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r0[i]) + (CAT ? __uint_as_float(r1[i]) : 0.f);
          if (KIND == 0) {
            if (b < a.B) {
              const int f0 = oc0 * 8 + c;
              const int bo = (a.mode == 2) ? f0 : c;        // conv: bias per channel; dense: per feature
#pragma unroll
              for (int i = 0; i < 16; i++) v[i] = act_apply(a.act, a.alpha * v[i] + a.bias[bo + i]);
              if (a.out.hi) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                  uint4 hi, lo;
                  split8(v + 8 * h, hi, lo);
                  bf16* p = a.out.hi + ((long long)(oc0 + (c >> 3) + h) * a.out.pitch + b) * 8;
                  *reinterpret_cast<uint4*>(p) = hi;
                  *reinterpret_cast<uint4*>(p + a.out.lo_off) = lo;
                }
              }
              if (a.out_f32) {
                float4* o = reinterpret_cast<float4*>(a.out_f32 + (long long)b * a.ld_f32 + f0);
#pragma unroll
                for (int i = 0; i < 4; i++) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              }
            }
          } else if (KIND == 1) {
            if (b < a.B) {
              float4* o = reinterpret_cast<float4*>(a.part + (long long)z * a.part_z + (long long)b * a.ld_part + oc0 * 8 + c);
#pragma unroll
              for (int i = 0; i < 4; i++) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          } else {
            if (b < a.B) {
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const long long e = ((long long)(oc0 + (c >> 3) + h) * a.out.pitch + b) * 8;
                float* vv = v + 8 * h;
                if (a.src_act == 1) {
                  float s[8];
                  unpack8(*reinterpret_cast<const uint4*>(a.src.hi + ((long long)(oc0 + (c >> 3) + h) * a.src.pitch + b) * 8), s);
#pragma unroll
                  for (int i = 0; i < 8; i++) vv[i] = s[i] > 0.f ? vv[i] : 0.f;
                } else if (a.src_act == 2) {
                  float s[8], t[8];
                  const bf16* sp = a.src.hi + ((long long)(oc0 + (c >> 3) + h) * a.src.pitch + b) * 8;
                  unpack8(*reinterpret_cast<const uint4*>(sp), s);
                  unpack8(*reinterpret_cast<const uint4*>(sp + a.src.lo_off), t);
#pragma unroll
                  for (int i = 0; i < 8; i++) { float y = s[i] + t[i]; vv[i] *= 1.f - y * y; }
                }
#pragma unroll
                for (int i = 0; i < 8; i++) dbacc[c + 8 * h + i] += vv[i];
                if (a.accumulate) {
                  float s[8], t[8];
                  unpack8(*reinterpret_cast<const uint4*>(a.out.hi + e), s);
                  unpack8(*reinterpret_cast<const uint4*>(a.out.hi + a.out.lo_off + e), t);
#pragma unroll
                  for (int i = 0; i < 8; i++) vv[i] += s[i] + t[i];
                }
                uint4 hi, lo;
                split8(vv, hi, lo);
                *reinterpret_cast<uint4*>(a.out.hi + e) = hi;
                *reinterpret_cast<uint4*>(a.out.hi + a.out.lo_off + e) = lo;
              }
            } else if (b < b_pad && !a.accumulate) {
              // gradient rows up to the next multiple of 16 are read by the weight-gradient K loop: keep them zero
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const long long e = ((long long)(oc0 + (c >> 3) + h) * a.out.pitch + b) * 8;
                *reinterpret_cast<uint4*>(a.out.hi + e) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(a.out.hi + a.out.lo_off + e) = make_uint4(0, 0, 0, 0);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
    if (KIND == 2 && a.db_part) {
      // column sums of this CTA: lanes -> warps -> one row of the partial table (added in order by grad_reduce_kernel)
#pragma unroll
      for (int j = 0; j < 64; j++) {
        if (j < N) {
          float s = dbacc[j];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0) red_sh[q][j] = s;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int t = tid - 64;
      if (t < N) a.db_part[(long long)blockIdx.x * N + t] = (red_sh[0][t] + red_sh[1][t]) + (red_sh[2][t] + red_sh[3][t]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// bp_wgrad_kernel: dW[k, n] = sum_b X[b, k] G[b, n].  Both operands are MN-major (their shared-memory rows are the
// reduction axis = samples), one K chunk = 64 samples.  M = 128 features of X = 16 consecutive feature chunks:
//   conv  : the KW*C features of one filter row at one output pixel are contiguous ("run"); accumulator r =
//           (filter row ky, M tile mt) holds dW rows (ky, kx, ci); a CTA walks a subset of (output pixel, 64 samples)
//           chunks, adds every filter row of each into its own TMEM accumulator and finally writes its partial sums
//           part[cta][r][128][N]; chunks that fall outside the image (SAME padding) are read from a zero buffer
//   dense : accumulator = (M tile r of the input features, N tile); the CTA reduces over all samples and stores dW.
// ------------------------------------------------------------------------------------------
constexpr int WG_KB = 64;                        // samples per K chunk
constexpr int WG_PLANE = WG_KB * 16;             // 1024
constexpr int WG_STAGE_A = 2 * 16 * WG_PLANE;    // hi + lo, 16 chunks
constexpr int WG_STAGE_B = 2 * 8 * WG_PLANE;     // hi + lo, up to 8 chunks (N <= 64)
constexpr int WG_STAGE = WG_STAGE_A + WG_STAGE_B;
constexpr int WG_STAGES = 4;

struct WgradArgs {
  BpT x; int x_split; BpT g; const bf16* zeros;
  int mode;                                      // 0 conv, 1 dense
  int H, W, C, OH, OW, KH, KW, S, padT, padL;
  int run_chunks, mts, R;                        // conv: chunks per filter row, M tiles per row, accumulators
  int N, B, n_bsub, n_opix;
  int x_chunks, r_tiles, n_ntiles;               // dense: feature chunks of X, M tiles, N tiles
  float* part;                                   // conv: partial sums [cta][R][128][N]
  float* dw; int ldw; int k_rows;                // dense: dW[k_rows][ldw]
};

__global__ void __launch_bounds__(RW_THREADS, 1)
bp_wgrad_kernel(const __grid_constant__ WgradArgs a, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ __align__(8) uint64_t bars[2 * WG_STAGES + 2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[WG_STAGES]);
  const uint32_t tfull = smem_u32(&bars[2 * WG_STAGES]), tempty = smem_u32(&bars[2 * WG_STAGES + 1]);
  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tfull, 1); mbar_init(tempty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t stage0 = smem_u32(smem);
  const int N = a.N;
  // work: conv = one pass over this CTA's (pixel, sample chunk) list; dense = tiles (r, ntile), each over all samples
  const int n_tiles = a.mode == 0 ? 1 : a.r_tiles * a.n_ntiles;
  const int n_kc = a.mode == 0 ? a.n_opix * a.n_bsub : a.n_bsub;

  if (warp == 0) {
    // ================= TMA producer =================
    int stage = 0; uint32_t phase = 0;
    for (int tile = (a.mode == 0 ? 0 : blockIdx.x); tile < n_tiles; tile += (a.mode == 0 ? 1 : gridDim.x)) {
      for (int kc = (a.mode == 0 ? blockIdx.x : 0); kc < n_kc; kc += (a.mode == 0 ? gridDim.x : 1)) {
        int opix = 0, bs = kc, rt = 0, nt = 0;
        if (a.mode == 0) { opix = kc / a.n_bsub; bs = kc - opix * a.n_bsub; }
        else { rt = tile / a.n_ntiles; nt = tile - rt * a.n_ntiles; }
        const int b0 = bs * WG_KB;
        const int nr16 = (min(WG_KB, a.B - b0) + 15) & ~15;
        const uint32_t bytes = (uint32_t)nr16 * 16;
        const int oy = opix / a.OW, ox = opix - oy * a.OW;
        const int n_r = a.mode == 0 ? a.R : 1;
        for (int r = 0; r < n_r; r++) {
          int chunk0, xc0 = 0, mt = 0;          // first X feature chunk of the 16-chunk M tile
          if (a.mode == 0) {
            const int ky = r / a.mts; mt = r - ky * a.mts;
            const int iy = oy * a.S - a.padT + ky;
            if (iy < 0 || iy >= a.H) continue;
            xc0 = ox * a.S - a.padL;
            chunk0 = ((iy * a.W + xc0) * a.C) / 8 + mt * 16;     // may be negative at a padded border: per-chunk test below
          } else chunk0 = rt * 16;
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const int n_a = 16 * (a.x_split ? 2 : 1), n_bh = N >> 3;
          if (lane == 0) mbar_expect_tx(full0 + 8 * stage, bytes * (n_a + 2 * n_bh));
          __syncwarp();
          const uint32_t sA = stage0 + (uint32_t)stage * WG_STAGE, sB = sA + WG_STAGE_A;
          for (int i = lane; i < n_a + 2 * n_bh; i += 32) {
            if (i < n_a) {
              const int pl = i >= 16, c = i - pl * 16;
              bool ok;
              if (a.mode == 0) {
                const int cj = mt * 16 + c;                       // chunk within the filter-row run
                const int px = xc0 + (cj * 8) / a.C;              // input column of that chunk
                ok = cj < a.run_chunks && px >= 0 && px < a.W;
              } else ok = chunk0 + c < a.x_chunks;
              const bf16* src = ok ? a.x.hi + (pl ? a.x.lo_off : 0) + ((long long)(chunk0 + c) * a.x.pitch + b0) * 8 : a.zeros;
              bulk_g2s(sA + pl * (16 * WG_PLANE) + c * WG_PLANE, src, bytes, full0 + 8 * stage);
            } else {
              const int j = i - n_a, pl = j >= n_bh, c = j - pl * n_bh;
              const int gc = (a.mode == 0 ? opix * (N >> 3) : nt * (N >> 3)) + c;
              const bf16* src = a.g.hi + (pl ? a.g.lo_off : 0) + ((long long)gc * a.g.pitch + b0) * 8;
              bulk_g2s(sB + pl * (8 * WG_PLANE) + c * WG_PLANE, src, bytes, full0 + 8 * stage);
            }
          }
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc(N, true, true);
    int stage = 0; uint32_t phase = 0; int cnt = 0;
    for (int tile = (a.mode == 0 ? 0 : blockIdx.x); tile < n_tiles; tile += (a.mode == 0 ? 1 : gridDim.x), cnt++) {
      mbar_wait(tempty, (cnt & 1) ^ 1);
      tc_fence_after();
      uint32_t touched = 0;
      for (int kc = (a.mode == 0 ? blockIdx.x : 0); kc < n_kc; kc += (a.mode == 0 ? gridDim.x : 1)) {
        int opix = 0, bs = kc;
        if (a.mode == 0) { opix = kc / a.n_bsub; bs = kc - opix * a.n_bsub; }
        const int b0 = bs * WG_KB;
        const int ksteps = ((min(WG_KB, a.B - b0) + 15) & ~15) >> 4;
        const int oy = opix / a.OW;
        const int n_r = a.mode == 0 ? a.R : 1;
        for (int r = 0; r < n_r; r++) {
          if (a.mode == 0) {
            const int iy = oy * a.S - a.padT + r / a.mts;
            if (iy < 0 || iy >= a.H) continue;
          }
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sA = stage0 + (uint32_t)stage * WG_STAGE, sB = sA + WG_STAGE_A;
            const uint32_t acc = tmem_base + r * N;
            uint32_t accum = (touched >> r) & 1u;
            for (int j = 0; j < ksteps; j++) {
              const uint64_t ah = make_desc(sA + j * 256, 128, WG_PLANE);
              const uint64_t gh = make_desc(sB + j * 256, 128, WG_PLANE);
              const uint64_t gl = make_desc(sB + 8 * WG_PLANE + j * 256, 128, WG_PLANE);
              umma_bf16(acc, ah, gh, idesc, accum);
              umma_bf16(acc, ah, gl, idesc, 1u);
              if (a.x_split) umma_bf16(acc, make_desc(sA + 16 * WG_PLANE + j * 256, 128, WG_PLANE), gh, idesc, 1u);
              accum = 1u;
            }
            umma_commit(empty0 + 8 * stage);
          }
          touched |= 1u << r;
          __syncwarp();
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (lane == 0) umma_commit(tfull);
      __syncwarp();
    }
  } else {
    // ================= epilogue =================
    const int q = warp & 3;
    const int m = q * 32 + lane;                 // accumulator row = feature within the M tile
    int cnt = 0;
    for (int tile = (a.mode == 0 ? 0 : blockIdx.x); tile < n_tiles; tile += (a.mode == 0 ? 1 : gridDim.x), cnt++) {
      mbar_wait(tfull, cnt & 1);
      tc_fence_after();
      // accumulators this CTA's chunk list reached (same walk as the producer / MMA roles)
      uint32_t touched = a.mode == 0 ? 0u : 1u;
      if (a.mode == 0) {
        for (int kc = blockIdx.x; kc < n_kc; kc += gridDim.x) {
          const int oy = (kc / a.n_bsub) / a.OW;
          for (int r = 0; r < a.R; r++) {
            const int iy = oy * a.S - a.padT + r / a.mts;
            if (iy >= 0 && iy < a.H) touched |= 1u << r;
          }
        }
      }
      const int n_r = a.mode == 0 ? a.R : 1;
      const bool have_work = a.mode != 0 || (int)blockIdx.x < a.n_opix * a.n_bsub;
      for (int r = 0; r < n_r; r++) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + r * N;
        float* dst;
        bool ok;
        if (a.mode == 0) {
          dst = a.part + (((long long)blockIdx.x * a.R + r) * 128 + m) * N;
          ok = have_work;
        } else {
          const int rt = tile / a.n_ntiles, nt = tile - rt * a.n_ntiles;
          const int k = rt * 128 + m;
          dst = a.dw + (long long)k * a.ldw + nt * N;
          ok = k < a.k_rows;
        }
        const bool live = (touched >> r) & 1u;
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          if (c < N) {
            uint32_t rr[16];
            tmem_ld16_nowait(taddr + c, rr);
            tmem_wait_ld();
            if (ok) {
              float4* o = reinterpret_cast<float4*>(dst + c);
#pragma unroll
              for (int i = 0; i < 4; i++)
                o[i] = live ? make_float4(__uint_as_float(rr[4 * i]), __uint_as_float(rr[4 * i + 1]), __uint_as_float(rr[4 * i + 2]),
                                          __uint_as_float(rr[4 * i + 3]))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// Ordered reduction of per-CTA partial sums into the flat gradient bucket (assignment, not accumulation).
// ------------------------------------------------------------------------------------------
struct RedSeg {
  const float* part; int n_slabs; long long slab;   // slab stride (floats); element idx of the segment inside a slab
  int count;                                         // elements of one slab that this segment covers
  int kind;                                          // 0: conv weight partials [R][128][N]; 1: plain vector (bias)
  int N, C, KW, mts, s2d_k4;                         // conv mapping
  long long dst_off; float alpha;
};
// real HWIO row of row k' = ((ty*k4 + tx)*16 + dy*4 + dx)*4 + c of the (k4 x k4, stride 1, 64 channel) conv over a
// space-to-depth plane: ((4ty+dy)*4k4 + 4tx+dx)*4 + c of the (4k4 x 4k4, stride 4, 4 channel) conv
__host__ __device__ inline int s2d_real_row(int m, int k4) {
  int tap = m >> 6, ty = tap / k4, tx = tap - ty * k4;
  int dy = (m >> 4) & 3, dx = (m >> 2) & 3, c = m & 3;
  return (((4 * ty + dy) * 4 * k4 + 4 * tx + dx) << 2) + c;
}
constexpr int RED_MAX = 24;
struct RedSegs { RedSeg s[RED_MAX]; };
__global__ void grad_reduce_kernel(const __grid_constant__ RedSegs segs, float* __restrict__ grads) {
  pdl_wait(); pdl_trigger();
  const RedSeg& s = segs.s[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.count) return;
  long long dst;
  if (s.kind == 0) {
    const int n = i % s.N, mm = (i / s.N) & 127, r = i / (s.N * 128);
    const int ky = r / s.mts, mt = r - ky * s.mts;
    const int feat = mt * 128 + mm;
    if (feat >= s.KW * s.C) return;
    int row = ky * s.KW * s.C + feat;
    if (s.s2d_k4) row = s2d_real_row(row, s.s2d_k4);
    dst = s.dst_off + (long long)row * s.N + n;
  } else dst = s.dst_off + i;
  float acc = 0.f;
  for (int k = 0; k < s.n_slabs; k++) acc += s.part[(long long)k * s.slab + i];
  grads[dst] = s.alpha * acc;
}

// ------------------------------------------------------------------------------------------
// layout kernels
// ------------------------------------------------------------------------------------------
// fp32 row-major [B][F] -> batch-planar hi/lo planes; rows [B, round16(B)) are zero-filled
__global__ void bp_split_kernel(const float* __restrict__ src, int B, int F, BpT dst) {
  pdl_wait(); pdl_trigger();
  const int b_pad = (B + 15) & ~15;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunks = F >> 3;
  if (i >= (long long)chunks * b_pad) return;
  const int c = (int)(i / b_pad), b = (int)(i - (long long)c * b_pad);
  uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
  if (b < B) {
    const float4* p = reinterpret_cast<const float4*>(src + (long long)b * F + c * 8);
    float4 x = p[0], y = p[1];
    float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
    split8(v, hi, lo);
  }
  bf16* o = dst.hi + ((long long)c * dst.pitch + b) * 8;
  *reinterpret_cast<uint4*>(o) = hi;
  *reinterpret_cast<uint4*>(o + dst.lo_off) = lo;
}
// batch-planar planes -> fp32 row-major [B][F]  (x = hi + lo)
__global__ void bp_merge_kernel(BpT src, int B, int F, float* __restrict__ dst) {
  pdl_wait(); pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunks = F >> 3;
  if (i >= (long long)chunks * B) return;
  const int c = (int)(i / B), b = (int)(i - (long long)c * B);
  const bf16* p = src.hi + ((long long)c * src.pitch + b) * 8;
  float h[8], l[8];
  unpack8(*reinterpret_cast<const uint4*>(p), h);
  unpack8(*reinterpret_cast<const uint4*>(p + src.lo_off), l);
  float4* o = reinterpret_cast<float4*>(dst + (long long)b * F + c * 8);
  o[0] = make_float4(h[0] + l[0], h[1] + l[1], h[2] + l[2], h[3] + l[3]);
  o[1] = make_float4(h[4] + l[4], h[5] + l[5], h[6] + l[6], h[7] + l[7]);
}
// column sums of a batch-planar gradient tensor (bias gradient when no data-gradient epilogue produced it):
// db[f % period] += sum_b g[b, f]; one warp per feature chunk
__global__ void bp_colsum_kernel(BpT g, int B, int F, int period, float* __restrict__ db) {
  pdl_wait(); pdl_trigger();
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= (F >> 3)) return;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int b = lane; b < B; b += 32) {
    const bf16* p = g.hi + ((long long)c * g.pitch + b) * 8;
    float h[8], l[8];
    unpack8(*reinterpret_cast<const uint4*>(p), h);
    unpack8(*reinterpret_cast<const uint4*>(p + g.lo_off), l);
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] += h[i] + l[i];
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float v = s[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) atomicAdd(db + (c * 8 + i) % period, v);
  }
}
// split-K partial sums part[z][b][N] -> act(sum_z + bias): fp32 row-major and/or batch-planar planes (fixed order)
__global__ void bp_splitk_finish_kernel(const float* __restrict__ part, int n_z, long long part_z, int B, int N,
                                        const float* __restrict__ bias, int act, float* __restrict__ out_f32, BpT out) {
  pdl_wait(); pdl_trigger();
  const int b_pad = (B + 15) & ~15;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunks = N >> 3;
  if (i >= (long long)chunks * b_pad) return;
  const int c = (int)(i / b_pad), b = (int)(i - (long long)c * b_pad);
  if (b >= B) return;
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int z = 0; z < n_z; z++) {
    const float4* p = reinterpret_cast<const float4*>(part + (long long)z * part_z + (long long)b * N + c * 8);
    float4 x = p[0], y = p[1];
    v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w; v[4] += y.x; v[5] += y.y; v[6] += y.z; v[7] += y.w;
  }
#pragma unroll
  for (int k = 0; k < 8; k++) v[k] = act_apply(act, v[k] + bias[c * 8 + k]);
  if (out_f32) {
    float4* o = reinterpret_cast<float4*>(out_f32 + (long long)b * N + c * 8);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (out.hi) {
    uint4 hi, lo;
    split8(v, hi, lo);
    bf16* o = out.hi + ((long long)c * out.pitch + b) * 8;
    *reinterpret_cast<uint4*>(o) = hi;
    *reinterpret_cast<uint4*>(o + out.lo_off) = lo;
  }
}

// Weight blobs: for every tensor-core layer the kernel matrix W[K, N] as batch-planar W^T planes
// blob[((n >> 3) * K + k') * 8 + (n & 7)], k' = space-to-depth row order for a stride-4 first layer.
struct BlobSeg { long long w_off, blob_off; int K, N, s2d_k4; };
__global__ void bp_wprep_kernel(const float* __restrict__ params, const BlobSeg* __restrict__ segs, bf16* __restrict__ hi,
                                long long lo_off) {
  pdl_wait(); pdl_trigger();
  const BlobSeg s = segs[blockIdx.y];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one (n chunk, k') piece per thread
  const int chunks = s.N >> 3;
  if (i >= (long long)chunks * s.K) return;
  const int c = (int)(i / s.K), kp = (int)(i - (long long)c * s.K);
  const int k = s.s2d_k4 ? s2d_real_row(kp, s.s2d_k4) : kp;
  const float4* p = reinterpret_cast<const float4*>(params + s.w_off + (long long)k * s.N + c * 8);
  float4 x = p[0], y = p[1];
  float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
  uint4 h, l;
  split8(v, h, l);
  bf16* o = hi + s.blob_off + ((long long)c * s.K + kp) * 8;
  *reinterpret_cast<uint4*>(o) = h;
  *reinterpret_cast<uint4*>(o + lo_off) = l;
}

// uint8 frames [n][H][W][4] (+ minibatch gather) -> batch-planar space-to-depth plane (exact in bf16, hi only):
// the image is embedded at (padT, padL) in a zero canvas of [4*H4][4*W4] pixels, pixel block (Y, X) = 64 features
// (dy, dx, c); a 16-byte output piece = 8 source bytes (2 pixels x 4 channels).
// One block = one block-row Y of 32 samples.  Phase 1 reads the 4 image rows of every sample with 16-byte loads
// along the row (contiguous 1344-byte runs per sample) into shared memory; phase 2 hands every piece to one warp
// whose lanes are the 32 samples, so each store instruction writes 512 contiguous bytes of the plane.
constexpr int DEC_SAMPLES = 32;
constexpr int DEC_THREADS = 256;
__global__ void __launch_bounds__(DEC_THREADS)
bp_decode_s2d_kernel(const uint8_t* __restrict__ obs, const int32_t* __restrict__ idx, int B, int H, int W, int H4, int W4,
                     int padT, int padL, BpT dst) {
  extern __shared__ __align__(16) uint8_t dec_sm[];
  pdl_wait(); pdl_trigger();
  const int Y = blockIdx.x, b0 = blockIdx.y * DEC_SAMPLES;
  const int row_bytes = W * 4, units = row_bytes >> 4;          // 16-byte units per image row
  const int sstride = 4 * row_bytes + 8;                        // per-sample stride: +8 B keeps 8-byte lane reads conflict-free
  const int ns = min(DEC_SAMPLES, B - b0);
  const long long img = (long long)H * row_bytes;
  for (int i = threadIdx.x; i < ns * 4 * units; i += DEC_THREADS) {
    const int sidx = i / (4 * units), rem = i - sidx * 4 * units, r = rem / units, ux = rem - r * units;
    const int y = 4 * Y + r - padT;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < H) {
      const long long srow = idx ? idx[b0 + sidx] : b0 + sidx;
      v = *reinterpret_cast<const uint4*>(obs + srow * img + (long long)y * row_bytes + ux * 16);
    }
    uint2* d = reinterpret_cast<uint2*>(dec_sm + sidx * sstride + r * row_bytes + ux * 16);
    d[0] = make_uint2(v.x, v.y);
    d[1] = make_uint2(v.z, v.w);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half_units = W >> 1;                                // 8-byte units (pixel pairs) per image row
  for (int q = warp; q < W4 * 8; q += DEC_THREADS / 32) {
    const int X = q >> 3, dy = (q >> 1) & 3, dxh = q & 1;
    const int u8 = 2 * X + dxh - (padL >> 1);                   // pixel pair inside the source row
    uint4 o = make_uint4(0, 0, 0, 0);
    if (lane < ns && u8 >= 0 && u8 < half_units) {
      const uint2 v = *reinterpret_cast<const uint2*>(dec_sm + lane * sstride + dy * row_bytes + u8 * 8);
      o.x = pack_bf16((float)(v.x & 0xff), (float)((v.x >> 8) & 0xff));
      o.y = pack_bf16((float)((v.x >> 16) & 0xff), (float)(v.x >> 24));
      o.z = pack_bf16((float)(v.y & 0xff), (float)((v.y >> 8) & 0xff));
      o.w = pack_bf16((float)((v.y >> 16) & 0xff), (float)(v.y >> 24));
    }
    if (lane < ns)
      *reinterpret_cast<uint4*>(dst.hi + ((long long)((Y * W4 + X) * 8 + dy * 2 + dxh) * dst.pitch + b0 + lane) * 8) = o;
  }
}

}  // namespace bp
}  // namespace xtb
