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
// Kernels (all persistent, one CTA per SM, warp specialised with replicated roles: 4 bulk-copy producer warps, each
// owning the ring slots s == warp (mod 4); 4 MMA-issue warps, each owning one in-flight output tile with its own TMEM
// accumulator and its own set of full barriers; 8 epilogue warps, two per TMEM lane quarter; every issue is done by one
// elected lane with warp-uniform operands; stage / unit tables and the conv weight blobs live in shared memory):
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

#ifndef XTB_BP_WAIT_HINT
#define XTB_BP_WAIT_HINT 20000u   // suspend-time hint (ns) of mbarrier.try_wait
#endif

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
      : "r"(bar), "r"(parity), "r"(XTB_BP_WAIT_HINT)
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
// one lane of a converged warp: the uniform-datapath instructions (UBLKCP, UTCHMMA, UTCBAR) are issued under it with
// warp-uniform operands, so ptxas keeps addresses and descriptors in uniform registers instead of wrapping every
// instruction in an ELECT / R2UR sequence
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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

// Optional timeline trace (XTB_BP_TRACE builds only): %globaltimer stamps of one lane per role, kept in shared memory
// while the kernel runs (a stamp is a timer read and one st.shared) and written out by the first CTAs at kernel exit.
#ifdef XTB_BP_TRACE
__device__ unsigned long long* g_bp_trace = nullptr;     // [cta < 4][role 0..2][BP_TR_N] ; slot 0 = count
#define BP_TR_N 22
#define BP_TRACE_DECL __shared__ unsigned long long bp_tr[3][BP_TR_N]; if (threadIdx.x < 3) bp_tr[threadIdx.x][0] = 0;
#define BP_TRACE(role, code) do { if ((threadIdx.x & 31) == 0 && ((role) != 0 || threadIdx.x < 32) && ((role) != 1 || threadIdx.x < 32 * (RW_P_WARPS + 1))) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); \
    unsigned long long i_ = bp_tr[role][0] + 1; if (i_ < BP_TR_N) { bp_tr[role][i_] = (t_ << 8) | (unsigned)(code); bp_tr[role][0] = i_; } } } while (0)
#define BP_TRACE_DUMP do { __syncthreads(); if (g_bp_trace && blockIdx.x < 4) for (int i_ = threadIdx.x; i_ < 3 * BP_TR_N; i_ += blockDim.x) \
    g_bp_trace[(size_t)blockIdx.x * 3 * BP_TR_N + i_] = bp_tr[i_ / BP_TR_N][i_ % BP_TR_N]; } while (0)
#else
#define BP_TRACE_DECL
#define BP_TRACE(role, code) do {} while (0)
#define BP_TRACE_DUMP do {} while (0)
#endif

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
//
// The stage walk of a conv unit (which feature chunks, which weight rows) is a host-built table (StageEnt per stage,
// UnitEnt per unit, built once per layer in xtb_net_create): the role loops are a table fetch, an mbarrier wait and a
// handful of adds per stage.  (First version: the walk was computed in the loops; a stage cost ~0.85 us of dependent
// integer code in the producer and MMA threads while its data arrived in 0.16 us -- profiles/r2_role_timeline.md.)
// ------------------------------------------------------------------------------------------
// Warp roles.  A single warp needs ~0.3-0.6 us of dependent instructions per stage (table fetch, barrier wait, address
// and descriptor arithmetic, issue) while the data of a stage arrives in ~0.16 us, so the issue roles are replicated:
// producer warp p fills ring slot p (every slot has its own full/empty barrier pair), MMA warp m owns the tiles with
// (tile count & 3) == m and with them TMEM accumulator buffer m (MMAs of one accumulator stay in one warp's program
// order; the four warps work on the four in-flight ring slots like a systolic wave), two epilogue warps per TMEM lane
// quarter split the accumulator columns.
constexpr int RW_P_WARPS = 4;
constexpr int RW_M_WARPS = 4;
constexpr int RW_EPI_WARPS = 8;
constexpr int RW_E_WARP0 = RW_P_WARPS + RW_M_WARPS;      // 8: (warp & 3) of warps 8..15 covers every TMEM lane quarter twice
constexpr int RW_E_TID0 = 32 * RW_E_WARP0;
constexpr int RW_THREADS = 32 * (RW_E_WARP0 + RW_EPI_WARPS);
constexpr int RW_A_PLANE = 2048;                 // 128 rows x 16 B
constexpr int RW_STAGE_A = 2 * 8 * RW_A_PLANE;   // hi + lo, 8 chunks (64 K elements)
constexpr int RW_STAGE_B = 16384;
constexpr int RW_MAX_STAGES = 4;

struct StageEnt {            // one K stage of a conv unit (8 bytes; the tables are copied to shared memory at kernel start)
  uint32_t a_chunk;          // first operand feature chunk of the stage
  uint16_t w_row;            // first weight-blob row of the stage (forward: k row; data gradient: tap * Cin)
  uint16_t nch;              // feature chunks in this stage (even, <= 8)
};
typedef uint32_t UnitEnt;    // first stage index | stage count << 24
struct StageDesc { uint32_t a_off, w_off, nch; };   // byte offsets inside a plane / blob plane

struct RowsArgs {
  BpT a; int a_split;                               // operand rows (activations / gradients); a_split: lo plane is read
  const bf16* w_hi; const bf16* w_lo; int w_pitch;  // weight blob planes
  int w_res; int w_res_chunks;                      // blob resident in shared memory (conv): chunks to load per plane
  int mode;                                         // 0 conv forward, 1 conv data gradient, 2 dense
  const StageEnt* stages; const UnitEnt* units;     // conv stage walk (global; n_stage_ents / n_units entries)
  int n_stage_ents;
  int kchunks, kc_split, n_ntiles;                  // dense: K chunks in total / per split, N tiles
  int n_units, n_btiles, B, N;                      // tiles = n_units * n_btiles; N = accumulator columns per plane
  // forward epilogue
  BpT out; float* out_f32; int ld_f32; const float* bias; float alpha; int act;
  float* part; long long part_z; int ld_part;       // split-K partial sums part[z][b][ld_part]
  // data-gradient epilogue: out = acc * act'(src) (+ out); column sums into db_part[cta][N] when non-NULL
  BpT src; int src_act; int accumulate; float* db_part;
};

// per-tile stage source: conv = run of the shared-memory table, dense = arithmetic progression.  setup() copies the
// kernel parameters it needs into registers once per role (see the epilogue note on constant-bank reloads).
struct TileWalk {
  const uint2* stages_sm; const uint32_t* units_sm;
  uint32_t a_pstride, w_pitch; int mode, n_ntiles, kc_split, kchunks, N;
  const uint2* st; uint32_t a0, w0, da, dw; int ns, nch_last;
  __device__ __forceinline__ void setup(const RowsArgs& a, const uint2* s_sm, const uint32_t* u_sm) {
    stages_sm = s_sm; units_sm = u_sm;
    a_pstride = (uint32_t)a.a.pitch * 16u; w_pitch = (uint32_t)a.w_pitch;
    mode = a.mode; n_ntiles = a.n_ntiles; kc_split = a.kc_split; kchunks = a.kchunks; N = a.N;
  }
  template <int KIND>
  __device__ __forceinline__ void init(int u) {
    if (mode == 2) {
      const int nt = u % n_ntiles, z = u / n_ntiles;
      const int cbeg = z * kc_split, cend = min(kchunks, cbeg + kc_split);
      ns = (cend - cbeg + 7) >> 3;
      nch_last = cend - cbeg - 8 * (ns - 1);
      a0 = (uint32_t)cbeg * a_pstride; da = 8u * a_pstride;
      if (KIND == 2) { w0 = ((uint32_t)cbeg * w_pitch + (uint32_t)nt * N) * 16u; dw = 8u * w_pitch * 16u; }
      else { w0 = ((uint32_t)(nt * (N >> 3)) * w_pitch + (uint32_t)cbeg * 8u) * 16u; dw = 1024u; }
      st = nullptr;
    } else {
      const uint32_t ue = units_sm[u];
      st = stages_sm + (ue & 0xffffffu); ns = (int)(ue >> 24);
    }
  }
  __device__ __forceinline__ StageDesc get(int s) const {
    if (st) {
      const uint2 v = st[s];
      return StageDesc{v.x * a_pstride, (v.y & 0xffffu) * 16u, v.y >> 16};
    }
    return StageDesc{a0 + (uint32_t)s * da, w0 + (uint32_t)s * dw, (uint32_t)(s == ns - 1 ? nch_last : 8)};
  }
};

template <int KIND>
__global__ void __launch_bounds__(RW_THREADS, 1)
bp_rows_kernel(const __grid_constant__ RowsArgs a, int n_stages, int stage_bytes, int wres_bytes, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  // stage-walk tables live behind the stage ring: [stages : n_stage_ents x 8 B][units : n_units x 4 B]
  uint2* stages_sm = reinterpret_cast<uint2*>(smem + wres_bytes + n_stages * stage_bytes);
  uint32_t* units_sm = reinterpret_cast<uint32_t*>(stages_sm + a.n_stage_ents);
  // full barriers exist once per MMA warp: a slot is consumed alternately by the two warps, and a waiter that saw only
  // every other completion of one barrier could not tell its phase from the one two fills earlier
  __shared__ __align__(8) uint64_t bars[(RW_M_WARPS + 1) * RW_MAX_STAGES + 2 * RW_M_WARPS + 1];
  __shared__ uint32_t tmem_slot;
  __shared__ float red_sh[4][64];
  __shared__ float bias_sh[64];
  BP_TRACE_DECL

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[RW_M_WARPS * RW_MAX_STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[(RW_M_WARPS + 1) * RW_MAX_STAGES]), tempty0 = tfull0 + 8 * RW_M_WARPS;
  const uint32_t wbar = tempty0 + 8 * RW_M_WARPS;
  if (tid == 0) {
    for (int s = 0; s < RW_M_WARPS * RW_MAX_STAGES; s++) mbar_init(full0 + 8 * s, 1);
    for (int s = 0; s < RW_MAX_STAGES; s++) mbar_init(empty0 + 8 * s, 1);
    for (int b = 0; b < RW_M_WARPS; b++) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, RW_EPI_WARPS); }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (warp == RW_P_WARPS) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // Static inputs are fetched BEFORE the dependency wait, i.e. under the tail of the preceding kernel: the stage-walk
  // tables never change after bind, and the weight blob of a conv layer was written at least two kernels back (by the
  // blob refresh of the previous optimiser step; the layer's own operand producer sits in between), so it is complete
  // by the time the previous kernel has passed its own wait and released this one.
  if (a.mode != 2) {      // every thread helps, one barrier
    const uint2* gs = reinterpret_cast<const uint2*>(a.stages);
    for (int i = tid; i < a.n_stage_ents; i += RW_THREADS) stages_sm[i] = __ldg(gs + i);
    for (int i = tid; i < a.n_units; i += RW_THREADS) units_sm[i] = __ldg(a.units + i);
  }
  if (warp == 0 && a.w_res && elect_one()) {
    const uint32_t bytes = (uint32_t)a.w_res_chunks * a.w_pitch * 16;
    mbar_expect_tx(wbar, 2 * bytes);
    bulk_g2s(smem_u32(smem), a.w_hi, bytes, wbar);
    bulk_g2s(smem_u32(smem) + bytes, a.w_lo, bytes, wbar);
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t wres = smem_u32(smem);                  // resident weights: hi plane chunks then lo plane chunks
  const uint32_t stage0 = wres + (uint32_t)wres_bytes;
  const int total = a.n_units * a.n_btiles;
  const int N = a.N;
  constexpr bool CAT = (KIND != 2);                      // forward: one MMA over [W_hi | W_lo], accumulator 2N columns
  const int acc_cols = CAT ? 2 * N : N;
  const int nbt = a.n_btiles;
  // tile = blockIdx.x + i * gridDim.x  ->  (unit u, batch tile bt), advanced without divisions
  const int du = (int)gridDim.x / nbt, dbt = (int)gridDim.x - du * nbt;

  if (warp < RW_P_WARPS) {
    // ================= TMA producers: warp p fills ring slot p; converged warp, one elected lane issues the copies ====
    const bool a_split = a.a_split != 0, w_res = a.w_res != 0;
    const uint32_t a_pstride = (uint32_t)a.a.pitch * 16u, w_pstride = (uint32_t)a.w_pitch * 16u;
    const char* a_hi = reinterpret_cast<const char*>(a.a.hi);
    const char* a_lo = a_hi + a.a.lo_off * 2;
    const char* w_hi = reinterpret_cast<const char*>(a.w_hi);
    const char* w_lo = reinterpret_cast<const char*>(a.w_lo);
    const int n_bh_fwd = N >> 3;
    int stage = 0; uint32_t phase = 0; int cnt = 0;
    int u = (int)blockIdx.x / nbt, bt = (int)blockIdx.x - u * nbt;
    const int Bsz = a.B;
    TileWalk tw; tw.setup(a, stages_sm, units_sm);
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, cnt++) {
      const int b0 = bt * 128, nr = min(128, Bsz - b0);
      const uint32_t a_bytes = (uint32_t)nr * 16u, row_off = (uint32_t)b0 * 16u;
      tw.init<KIND>(u);
      for (int s = 0; s < tw.ns; s++) {
        if (stage != warp) { if (++stage == n_stages) { stage = 0; phase ^= 1; } continue; }
        const StageDesc d = tw.get(s);
        BP_TRACE(0, 1);
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        BP_TRACE(0, 2);
        const int nch = (int)d.nch;
        const int n_bh = w_res ? 0 : (KIND == 2 ? nch : n_bh_fwd);
        const uint32_t b_bytes = (KIND == 2) ? (uint32_t)N * 16u : (uint32_t)nch * 128u;
        const uint32_t fb = full0 + 8 * ((cnt & (RW_M_WARPS - 1)) * RW_MAX_STAGES + stage);
        if (elect_one()) {
#ifdef XTB_EXP_NO_LOAD
          mbar_arrive(fb);
#else
          mbar_expect_tx(fb, a_bytes * (a_split ? 2 * nch : nch) + b_bytes * 2 * n_bh);
          const uint32_t sA = stage0 + (uint32_t)stage * stage_bytes, sB = sA + RW_STAGE_A;
          const char* ah = a_hi + d.a_off + row_off;
          for (int c = 0; c < nch; c++) bulk_g2s(sA + c * RW_A_PLANE, ah + (uint32_t)c * a_pstride, a_bytes, fb);
          if (a_split) {
            const char* al = a_lo + d.a_off + row_off;
            for (int c = 0; c < nch; c++) bulk_g2s(sA + 8 * RW_A_PLANE + c * RW_A_PLANE, al + (uint32_t)c * a_pstride, a_bytes, fb);
          }
          if (!w_res) {
            const char* wh = w_hi + d.w_off; const char* wl = w_lo + d.w_off;
            if (KIND == 2) {
              for (int c = 0; c < n_bh; c++) {
                bulk_g2s(sB + c * (N * 16), wh + (uint32_t)c * w_pstride, b_bytes, fb);
                bulk_g2s(sB + 8192 + c * (N * 16), wl + (uint32_t)c * w_pstride, b_bytes, fb);
              }
            } else {
              for (int c = 0; c < n_bh; c++) {
                bulk_g2s(sB + c * 1024, wh + (uint32_t)c * w_pstride, b_bytes, fb);
                bulk_g2s(sB + (n_bh_fwd + c) * 1024, wl + (uint32_t)c * w_pstride, b_bytes, fb);
              }
            }
          }
#endif
        }
        __syncwarp();
        BP_TRACE(0, 3);
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
      u += du; bt += dbt;
      if (bt >= nbt) { bt -= nbt; u++; }
    }
  } else if (warp < RW_E_WARP0) {
    // ================= MMA issuers: warp m owns every second tile; converged warp, one elected lane issues ==========
    {
      const int mw = warp - RW_P_WARPS;
      const uint32_t idesc_n = make_idesc(N, false, !(KIND == 2));
      const uint32_t idesc_cat = make_idesc(2 * N, false, true);
      const bool a_split = a.a_split != 0, w_res = a.w_res != 0;
      // descriptor constant parts; the low 14 bits hold (shared address >> 4) and are advanced by plain adds
      const uint64_t adesc_hi = make_desc(0, RW_A_PLANE, 128);
      const uint64_t bdesc_hi = (KIND == 2) ? make_desc(0, (uint32_t)N * 16, 128) : make_desc(0, 128, 1024);
      const uint32_t a_lo_add = (8 * RW_A_PLANE) >> 4, a_step = (2 * RW_A_PLANE) >> 4;
      const uint32_t w_pstride = (uint32_t)a.w_pitch * 16u;
      // resident weights: forward MN-major (LBO 128, SBO = pitch rows), data gradient K-major (LBO = pitch rows, SBO 128)
      const uint64_t wdesc_hi = (KIND == 2) ? make_desc(0, w_pstride, 128) : make_desc(0, 128, w_pstride);
      const uint32_t w_lo_add = ((uint32_t)a.w_res_chunks * w_pstride) >> 4;
      const uint32_t b_step = (KIND == 2) ? (w_res ? (2 * w_pstride) >> 4 : (2u * N * 16u) >> 4) : (256 >> 4);
      const uint32_t b_lo_add = w_res ? w_lo_add : (8192 >> 4);
      const uint64_t b_const = w_res ? wdesc_hi : bdesc_hi;
      if (w_res) mbar_wait(wbar, 0);
      const uint32_t myfull = full0 + 8 * (mw * RW_MAX_STAGES);
      uint32_t fphase = 0;                 // bit s: parity of the next completion of this warp's full barrier of slot s
      int stage = 0; int cnt = 0;
      int u = (int)blockIdx.x / nbt, bt = (int)blockIdx.x - u * nbt;
      TileWalk tw; tw.setup(a, stages_sm, units_sm);
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, cnt++) {
        const int buf = cnt & (RW_M_WARPS - 1);
        const uint32_t acc = tmem_base + buf * acc_cols;
        tw.init<KIND>(u);
        if (buf != mw) {       // another warp's tile: only step the ring position over its stages
          stage = (stage + tw.ns) % n_stages;
          u += du; bt += dbt;
          if (bt >= nbt) { bt -= nbt; u++; }
          continue;
        }
        BP_TRACE(1, 10);
        mbar_wait(tempty0 + 8 * buf, ((cnt / RW_M_WARPS) & 1) ^ 1);
        tc_fence_after();
        BP_TRACE(1, 11);
        uint32_t accum = 0;
        for (int s = 0; s < tw.ns; s++) {
          const StageDesc d = tw.get(s);
          mbar_wait(myfull + 8 * stage, (fphase >> stage) & 1u);
          fphase ^= 1u << stage;
          tc_fence_after();
          BP_TRACE(1, 12);
          const uint32_t sA = stage0 + (uint32_t)stage * stage_bytes;
          uint32_t a_lo32 = sA >> 4;
          uint32_t b_lo32 = w_res ? (wres + d.w_off) >> 4 : (sA + RW_STAGE_A) >> 4;
          const int ksteps = (int)d.nch >> 1;
          if (elect_one()) {
#ifndef XTB_EXP_NO_MMA
            for (int j = 0; j < ksteps; j++) {
              const uint64_t ah = adesc_hi | a_lo32, bh = b_const | b_lo32;
              if (KIND != 2) {
                umma_bf16(acc, ah, bh, idesc_cat, accum);                                   // A_hi x [W_hi | W_lo]
                if (a_split) umma_bf16(acc, adesc_hi | (a_lo32 + a_lo_add), bh, idesc_n, 1u);   // A_lo x W_hi
              } else {
                umma_bf16(acc, ah, bh, idesc_n, accum);
                umma_bf16(acc, ah, b_const | (b_lo32 + b_lo_add), idesc_n, 1u);
                if (a_split) umma_bf16(acc, adesc_hi | (a_lo32 + a_lo_add), bh, idesc_n, 1u);
              }
              accum = 1u;
              a_lo32 += a_step; b_lo32 += b_step;
            }
#endif
            umma_commit(empty0 + 8 * stage);
          }
          accum = 1u;
          __syncwarp();
          BP_TRACE(1, 13);
          if (++stage == n_stages) stage = 0;
        }
        if (elect_one()) umma_commit(tfull0 + 8 * buf);
        __syncwarp();
        u += du; bt += dbt;
        if (bt >= nbt) { bt -= nbt; u++; }
      }
    }
  } else {
    // ================= epilogue: warps 2..9; warp & 3 = TMEM lane quarter, two warps per quarter split the columns ====
    // kernel parameters used per tile live in registers: the barrier waits are asm volatile with a memory clobber, after
    // which the compiler re-reads the constant bank (a chain of dependent ~50-cycle loads per tile otherwise)
    const int act = a.act, mode = a.mode, Bsz = a.B, n_nt = a.n_ntiles, ld_f32 = a.ld_f32, src_act = a.src_act, accumulate = a.accumulate;
    const int ld_part = a.ld_part;
    const float alpha = a.alpha;
    bf16* const out_hi = a.out.hi; const long long out_lo = a.out.lo_off;
    float* const out_f32 = a.out_f32; const float* const bias_g = a.bias;
    float* const part = a.part; const long long part_z = a.part_z;
    const bf16* const src_hi = a.src.hi; const long long src_lo = a.src.lo_off, src_pstride = (long long)a.src.pitch * 8;
    const int q = warp & 3, half = (warp - RW_E_WARP0) >> 2;
    if (KIND == 0 && mode != 2 && tid >= RW_E_TID0 && tid < RW_E_TID0 + N) bias_sh[tid - RW_E_TID0] = bias_g[tid - RW_E_TID0];
    asm volatile("bar.sync 1, %0;" ::"n"(32 * RW_EPI_WARPS) : "memory");
    float dbacc[32];
    if (KIND == 2) {
#pragma unroll
      for (int j = 0; j < 32; j++) dbacc[j] = 0.f;
    }
    const int b_pad = (Bsz + 15) & ~15;
    // columns of this warp: 16-column groups g with (g & 1) == half when N >= 32; the single group of N = 16 goes to half 0
    const int n_groups = N >> 4;
    const long long out_pstride = (long long)a.out.pitch * 8;      // elements between feature chunks
    int cnt = 0;
    int u = (int)blockIdx.x / nbt, bt = (int)blockIdx.x - u * nbt;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, cnt++) {
      const int buf = cnt & (RW_M_WARPS - 1);
      const int b = bt * 128 + q * 32 + lane;
      int oc0, z = 0;                          // first output chunk of the unit
      if (mode == 2) { const int nt = u % n_nt; z = u / n_nt; oc0 = nt * (N >> 3); }
      else oc0 = u * (N >> 3);
      if (warp == RW_E_WARP0) BP_TRACE(2, 20);
      mbar_wait(tfull0 + 8 * buf, (cnt / RW_M_WARPS) & 1);
      tc_fence_after();
      if (warp == RW_E_WARP0) BP_TRACE(2, 21);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * acc_cols;
      // all TMEM loads of this warp first, one wait
      uint32_t r0[2][16], r1[2][16];
#pragma unroll
      for (int gi = 0; gi < 2; gi++) {
        const int g = 2 * gi + half;
        if (g < n_groups) {
          tmem_ld16_nowait(taddr + 16 * g, r0[gi]);
          if (CAT) tmem_ld16_nowait(taddr + N + 16 * g, r1[gi]);
        }
      }
      tmem_wait_ld();
#ifndef XTB_EXP_NO_EPI
#pragma unroll
      for (int gi = 0; gi < 2; gi++) {
        const int g = 2 * gi + half;
        if (g < n_groups) {
          const int c = 16 * g;
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r0[gi][i]) + (CAT ? __uint_as_float(r1[gi][i]) : 0.f);
          if (KIND == 0) {
            if (b < Bsz) {
              const int f0 = oc0 * 8 + c;
              float bb[16];
              if (mode == 2) {
                const float4* bp4 = reinterpret_cast<const float4*>(bias_g + f0);
#pragma unroll
                for (int i = 0; i < 4; i++) { const float4 t4 = __ldg(bp4 + i); bb[4 * i] = t4.x; bb[4 * i + 1] = t4.y; bb[4 * i + 2] = t4.z; bb[4 * i + 3] = t4.w; }
              } else {
#pragma unroll
                for (int i = 0; i < 16; i++) bb[i] = bias_sh[c + i];
              }
              if (act == 2) {
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = tanhf(alpha * v[i] + bb[i]);
              } else {          // relu / linear without a branch per element
                const float lo = act == 1 ? 0.f : -INFINITY;
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = fmaxf(fmaf(alpha, v[i], bb[i]), lo);
              }
              if (out_hi) {
                bf16* p = out_hi + (long long)(oc0 + 2 * g) * out_pstride + (long long)b * 8;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                  uint4 hi, lo;
                  split8(v + 8 * h, hi, lo);
                  *reinterpret_cast<uint4*>(p + h * out_pstride) = hi;
                  *reinterpret_cast<uint4*>(p + h * out_pstride + out_lo) = lo;
                }
              }
              if (out_f32) {
                float4* o = reinterpret_cast<float4*>(out_f32 + (long long)b * ld_f32 + f0);
#pragma unroll
                for (int i = 0; i < 4; i++) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              }
            }
          } else if (KIND == 1) {
            if (b < Bsz) {
              float4* o = reinterpret_cast<float4*>(part + (long long)z * part_z + (long long)b * ld_part + oc0 * 8 + c);
#pragma unroll
              for (int i = 0; i < 4; i++) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          } else {
            bf16* p = out_hi + (long long)(oc0 + 2 * g) * out_pstride + (long long)b * 8;
            if (b < Bsz) {
              const bf16* sp = src_hi + (long long)(oc0 + 2 * g) * src_pstride + (long long)b * 8;
#pragma unroll
              for (int h = 0; h < 2; h++) {
                float* vv = v + 8 * h;
                if (src_act == 1) {
                  float s[8];
                  unpack8(*reinterpret_cast<const uint4*>(sp + h * src_pstride), s);
#pragma unroll
                  for (int i = 0; i < 8; i++) vv[i] = s[i] > 0.f ? vv[i] : 0.f;
                } else if (src_act == 2) {
                  float s[8], t[8];
                  unpack8(*reinterpret_cast<const uint4*>(sp + h * src_pstride), s);
                  unpack8(*reinterpret_cast<const uint4*>(sp + h * src_pstride + src_lo), t);
#pragma unroll
                  for (int i = 0; i < 8; i++) { float y = s[i] + t[i]; vv[i] *= 1.f - y * y; }
                }
#pragma unroll
                for (int i = 0; i < 8; i++) dbacc[16 * gi + 8 * h + i] += vv[i];
                if (accumulate) {
                  float s[8], t[8];
                  unpack8(*reinterpret_cast<const uint4*>(p + h * out_pstride), s);
                  unpack8(*reinterpret_cast<const uint4*>(p + h * out_pstride + out_lo), t);
#pragma unroll
                  for (int i = 0; i < 8; i++) vv[i] += s[i] + t[i];
                }
                uint4 hi, lo;
                split8(vv, hi, lo);
                *reinterpret_cast<uint4*>(p + h * out_pstride) = hi;
                *reinterpret_cast<uint4*>(p + h * out_pstride + out_lo) = lo;
              }
            } else if (b < b_pad && !accumulate) {
              // gradient rows up to the next multiple of 16 are read by the weight-gradient K loop: keep them zero
#pragma unroll
              for (int h = 0; h < 2; h++) {
                *reinterpret_cast<uint4*>(p + h * out_pstride) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(p + h * out_pstride + out_lo) = make_uint4(0, 0, 0, 0);
              }
            }
          }
        }
      }
#endif
      tc_fence_before();
      __syncwarp();
      if (warp == RW_E_WARP0) BP_TRACE(2, 22);
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
      u += du; bt += dbt;
      if (bt >= nbt) { bt -= nbt; u++; }
    }
    if (KIND == 2 && a.db_part) {
      // column sums of this CTA: lanes -> the four quarter warps of each column owner -> one row of the partial table
#pragma unroll
      for (int gi = 0; gi < 2; gi++) {
        const int g = 2 * gi + half;
        if (g < n_groups) {
#pragma unroll
          for (int i = 0; i < 16; i++) {
            float s = dbacc[16 * gi + i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red_sh[q][16 * g + i] = s;
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * RW_EPI_WARPS) : "memory");
      const int t = tid - RW_E_TID0;
      if (t < N) a.db_part[(long long)blockIdx.x * N + t] = (red_sh[0][t] + red_sh[1][t]) + (red_sh[2][t] + red_sh[3][t]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RW_P_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
  BP_TRACE_DUMP;
}

// ------------------------------------------------------------------------------------------
// bp_wgrad_kernel: dW[k, n] = sum_b X[b, k] G[b, n].  Both operands are MN-major (their shared-memory rows are the
// reduction axis = samples), one K chunk = 64 samples.  M = 128 features of X = 16 consecutive feature chunks:
//   conv  : the KW*C features of one filter row at one output pixel are contiguous ("run"); accumulator r =
//           (filter row ky, M tile mt) holds dW rows (ky, kx, ci); a CTA walks a subset of (output pixel, 64 samples)
//           chunks, adds every filter row of each into its own TMEM accumulator and finally writes its partial sums
//           part[cta][r][128][N]; chunks that fall outside the image (SAME padding) are read from a zero buffer.
//           The (pixel, accumulator) -> operand mapping is a host-built table (WgEnt).
//   dense : accumulator = (M tile r of the input features, N tile); the CTA reduces over all samples and stores dW.
// ------------------------------------------------------------------------------------------
constexpr int WG_KB = 64;                        // samples per K chunk
constexpr int WG_PLANE = WG_KB * 16;             // 1024
constexpr int WG_STAGE_A = 2 * 16 * WG_PLANE;    // hi + lo, 16 chunks
constexpr int WG_STAGE_B = 2 * 8 * WG_PLANE;     // hi + lo, up to 8 chunks (N <= 64)
constexpr int WG_STAGE = WG_STAGE_A + WG_STAGE_B;
constexpr int WG_STAGES = 4;
constexpr int WG_P_WARPS = 4;                   // producer warp p fills ring slot p
constexpr int WG_M_WARPS = 4;                   // MMA warp m owns the accumulators with (r & 3) == m
constexpr int WG_E_WARP0 = WG_P_WARPS + WG_M_WARPS;
constexpr int WG_THREADS = 32 * (WG_E_WARP0 + 4);

struct WgEnt {               // conv: one (output pixel, accumulator) pair (8 bytes, copied to shared memory)
  int32_t x_chunk;           // first X feature chunk of the M tile (may be negative at a padded border)
  uint16_t okmask;           // bit c: chunk c of the tile lies inside the image and inside the filter row
  uint16_t valid;            // the filter row exists for this pixel
};

struct WgradArgs {
  BpT x; int x_split; BpT g; const bf16* zeros;
  int mode;                                      // 0 conv, 1 dense
  const WgEnt* tab;                              // conv: [n_opix][R]
  int R;                                         // conv: accumulators
  int N, B, n_bsub, n_opix;
  int x_chunks, r_tiles, n_ntiles;               // dense: feature chunks of X, M tiles, N tiles
  float* part;                                   // conv: partial sums [cta][R][128][N]
  float* dw; int ldw; int k_rows;                // dense: dW[k_rows][ldw]
};

// stage source of one (K chunk, accumulator) pair
struct WgWalk {
  __device__ __forceinline__ static WgEnt get(const WgradArgs& a, const uint2* tab_sm, int opix, int r, int rt) {
    if (a.mode == 0) {
      const uint2 v = tab_sm[opix * a.R + r];
      return WgEnt{(int32_t)v.x, (uint16_t)(v.y & 0xffffu), (uint16_t)(v.y >> 16)};
    }
    const int c0 = rt * 16, left = a.x_chunks - c0;
    return WgEnt{c0, (uint16_t)(left >= 16 ? 0xffffu : ((1u << left) - 1u)), (uint16_t)1};
  }
};

__global__ void __launch_bounds__(WG_THREADS, 1)
bp_wgrad_kernel(const __grid_constant__ WgradArgs a, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ __align__(8) uint64_t bars[(WG_M_WARPS + 1) * WG_STAGES + 2];      // full barriers once per MMA warp (see bp_rows_kernel)
  __shared__ uint32_t tmem_slot;
  uint2* tab_sm = reinterpret_cast<uint2*>(smem + WG_STAGES * WG_STAGE);     // conv: [n_opix][R] entries behind the stage ring

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[WG_M_WARPS * WG_STAGES]);
  const uint32_t tfull = smem_u32(&bars[(WG_M_WARPS + 1) * WG_STAGES]), tempty = tfull + 8;
  if (tid == 0) {
    for (int s = 0; s < WG_M_WARPS * WG_STAGES; s++) mbar_init(full0 + 8 * s, 1);
    for (int s = 0; s < WG_STAGES; s++) mbar_init(empty0 + 8 * s, 1);
    mbar_init(tfull, WG_M_WARPS); mbar_init(tempty, 4);
    fence_barrier_init();
  }
  if (warp == WG_P_WARPS) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)tmem_cols);
  const bool conv = a.mode == 0;
  if (conv) {             // static table: fetched under the tail of the preceding kernel
    const uint2* gt = reinterpret_cast<const uint2*>(a.tab);
    for (int i = tid; i < a.n_opix * a.R; i += WG_THREADS) tab_sm[i] = __ldg(gt + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t stage0 = smem_u32(smem);
  const int N = a.N;
  // work: conv = one pass over this CTA's (pixel, sample chunk) list; dense = tiles (r, ntile), each over all samples
  const int n_tiles = conv ? 1 : a.r_tiles * a.n_ntiles;
  const int n_kc = conv ? a.n_opix * a.n_bsub : a.n_bsub;
  const int tile0 = conv ? 0 : (int)blockIdx.x, tile_step = conv ? 1 : (int)gridDim.x;
  const int kc0 = conv ? (int)blockIdx.x : 0, kc_step = conv ? (int)gridDim.x : 1;
  const int n_r = conv ? a.R : 1;
  const int dop = kc_step / a.n_bsub, dbs = kc_step - dop * a.n_bsub;      // (pixel, sample chunk) advance of one kc step

  if (warp < WG_P_WARPS) {
    // ================= TMA producers: warp p fills ring slot p; converged warp, one elected lane issues the copies ====
    const uint32_t pstride_x = (uint32_t)a.x.pitch * 16u, pstride_g = (uint32_t)a.g.pitch * 16u;
    const char* x_hi = reinterpret_cast<const char*>(a.x.hi);
    const char* x_lo = x_hi + a.x.lo_off * 2;
    const char* g_hi = reinterpret_cast<const char*>(a.g.hi);
    const char* g_lo = g_hi + a.g.lo_off * 2;
    const char* zeros = reinterpret_cast<const char*>(a.zeros);
    const bool x_split = a.x_split != 0;
    const int n_bh = N >> 3;
    int stage = 0; uint32_t phase = 0;
    for (int tile = tile0; tile < n_tiles; tile += tile_step) {
      const int rt = tile / a.n_ntiles, nt = tile - rt * a.n_ntiles;
      int opix = conv ? kc0 / a.n_bsub : 0, bs = conv ? kc0 - opix * a.n_bsub : 0;
      for (int kc = kc0; kc < n_kc; kc += kc_step) {
        const int b0 = bs * WG_KB;
        const uint32_t bytes = (uint32_t)((min(WG_KB, a.B - b0) + 15) & ~15) * 16u, row_off = (uint32_t)b0 * 16u;
        const uint32_t g_off = (uint32_t)((conv ? opix : nt) * n_bh) * pstride_g + row_off;
        for (int r = 0; r < n_r; r++) {
          const WgEnt e = WgWalk::get(a, tab_sm, opix, r, rt);
          if (!e.valid) continue;
          if (stage != warp) { if (++stage == WG_STAGES) { stage = 0; phase ^= 1; } continue; }
          const long long x_off = (long long)e.x_chunk * (long long)pstride_x + row_off;
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t fb = full0 + 8 * ((r & (WG_M_WARPS - 1)) * WG_STAGES + stage);
          if (elect_one()) {
            mbar_expect_tx(fb, bytes * ((x_split ? 32 : 16) + 2 * n_bh));
            const uint32_t sA = stage0 + (uint32_t)stage * WG_STAGE, sB = sA + WG_STAGE_A;
            const uint32_t mask = e.okmask;
#pragma unroll 4
            for (int c = 0; c < 16; c++)
              bulk_g2s(sA + c * WG_PLANE, ((mask >> c) & 1u) ? x_hi + x_off + (long long)((uint32_t)c * pstride_x) : zeros, bytes, fb);
            if (x_split) {
#pragma unroll 4
              for (int c = 0; c < 16; c++)
                bulk_g2s(sA + (16 + c) * WG_PLANE, ((mask >> c) & 1u) ? x_lo + x_off + (long long)((uint32_t)c * pstride_x) : zeros, bytes, fb);
            }
            for (int c = 0; c < n_bh; c++) {
              bulk_g2s(sB + c * WG_PLANE, g_hi + g_off + (uint32_t)c * pstride_g, bytes, fb);
              bulk_g2s(sB + (8 + c) * WG_PLANE, g_lo + g_off + (uint32_t)c * pstride_g, bytes, fb);
            }
          }
          __syncwarp();
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        if (conv) { opix += dop; bs += dbs; if (bs >= a.n_bsub) { bs -= a.n_bsub; opix++; } } else bs++;
      }
    }
  } else if (warp < WG_E_WARP0) {
    // ================= MMA issuers: warp m owns the accumulators with (r & 1) == m ====================================
    {
      const int mw = warp - WG_P_WARPS;
      const uint32_t idesc = make_idesc(N, true, true);
      const bool x_split = a.x_split != 0;
      const uint64_t desc_hi = make_desc(0, 128, WG_PLANE);      // MN-major operands: LBO 128 B, SBO = chunk plane
      const uint32_t myfull = full0 + 8 * (mw * WG_STAGES);
      uint32_t fphase = 0;
      int stage = 0; int cnt = 0;
      for (int tile = tile0; tile < n_tiles; tile += tile_step, cnt++) {
        const int rt = tile / a.n_ntiles;
        mbar_wait(tempty, (cnt & 1) ^ 1);
        tc_fence_after();
        uint32_t touched = 0;
        int opix = conv ? kc0 / a.n_bsub : 0, bs = conv ? kc0 - opix * a.n_bsub : 0;
        for (int kc = kc0; kc < n_kc; kc += kc_step) {
          const int b0 = bs * WG_KB;
          const int ksteps = ((min(WG_KB, a.B - b0) + 15) & ~15) >> 4;
          for (int r = 0; r < n_r; r++) {
            const WgEnt e = WgWalk::get(a, tab_sm, opix, r, rt);
            if (!e.valid) continue;
            if ((r & (WG_M_WARPS - 1)) != mw) { if (++stage == WG_STAGES) stage = 0; continue; }
            mbar_wait(myfull + 8 * stage, (fphase >> stage) & 1u);
            fphase ^= 1u << stage;
            tc_fence_after();
            const uint32_t sA = stage0 + (uint32_t)stage * WG_STAGE;
            uint32_t a32 = sA >> 4, g32 = (sA + WG_STAGE_A) >> 4;
            const uint32_t acc = tmem_base + r * N;
            uint32_t accum = (touched >> r) & 1u;
            if (elect_one()) {
              for (int j = 0; j < ksteps; j++) {
                const uint64_t ah = desc_hi | a32, gh = desc_hi | g32;
                umma_bf16(acc, ah, gh, idesc, accum);
                umma_bf16(acc, ah, desc_hi | (g32 + ((8 * WG_PLANE) >> 4)), idesc, 1u);
                if (x_split) umma_bf16(acc, desc_hi | (a32 + ((16 * WG_PLANE) >> 4)), gh, idesc, 1u);
                accum = 1u;
                a32 += 256 >> 4; g32 += 256 >> 4;
              }
              umma_commit(empty0 + 8 * stage);
            }
            __syncwarp();
            touched |= 1u << r;
            if (++stage == WG_STAGES) stage = 0;
          }
          if (conv) { opix += dop; bs += dbs; if (bs >= a.n_bsub) { bs -= a.n_bsub; opix++; } } else bs++;
        }
        if (elect_one()) umma_commit(tfull);
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue =================
    const int q = warp & 3;
    const int m = q * 32 + lane;                 // accumulator row = feature within the M tile
    int cnt = 0;
    for (int tile = tile0; tile < n_tiles; tile += tile_step, cnt++) {
      // accumulators this CTA's chunk list reaches (same walk as the producer / MMA roles), found while the main loop runs
      uint32_t touched = conv ? 0u : 1u;
      if (conv) {
        int opix = kc0 / a.n_bsub, bs = kc0 - opix * a.n_bsub;
        for (int kc = kc0; kc < n_kc; kc += kc_step) {
          for (int r = 0; r < a.R; r++) if (WgWalk::get(a, tab_sm, opix, r, 0).valid) touched |= 1u << r;
          opix += dop; bs += dbs; if (bs >= a.n_bsub) { bs -= a.n_bsub; opix++; }
        }
      }
      mbar_wait(tfull, cnt & 1);
      tc_fence_after();
      const bool have_work = !conv || (int)blockIdx.x < n_kc;
      for (int r = 0; r < n_r; r++) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + r * N;
        float* dst;
        bool ok;
        if (conv) {
          dst = a.part + (((long long)blockIdx.x * a.R + r) * 128 + m) * N;
          ok = have_work;
        } else {
          const int rt = tile / a.n_ntiles, nt = tile - rt * a.n_ntiles;
          const int k = rt * 128 + m;
          dst = a.dw + (long long)k * a.ldw + nt * N;
          ok = k < a.k_rows;
        }
        const bool live = (touched >> r) & 1u;
        uint32_t rr[4][16];
#pragma unroll
        for (int g = 0; g < 4; g++) if (16 * g < N) tmem_ld16_nowait(taddr + 16 * g, rr[g]);
        tmem_wait_ld();
        if (ok) {
#pragma unroll
          for (int g = 0; g < 4; g++) {
            if (16 * g < N) {
              float4* o = reinterpret_cast<float4*>(dst + 16 * g);
#pragma unroll
              for (int i = 0; i < 4; i++)
                o[i] = live ? make_float4(__uint_as_float(rr[g][4 * i]), __uint_as_float(rr[g][4 * i + 1]), __uint_as_float(rr[g][4 * i + 2]),
                                          __uint_as_float(rr[g][4 * i + 3]))
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
  if (warp == WG_P_WARPS) {
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
  float* dst_ptr;                                    // destination base instead of the gradient bucket (kind 1), or NULL
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
// One block = 128 consecutive elements of one segment: warp w sums slabs w, w + 8, ... with float4 loads (512 B per
// warp and slab), then warp 0 adds the eight partial sums in warp order -- a fixed tree, bitwise reproducible.
__global__ void __launch_bounds__(256) grad_reduce_kernel(const __grid_constant__ RedSegs segs, float* __restrict__ grads) {
  pdl_wait(); pdl_trigger();
  __shared__ float4 sh[8][32];
  const RedSeg& s = segs.s[blockIdx.y];
  if ((int)blockIdx.x * 128 >= s.count) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = ((int)blockIdx.x * 32 + lane) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < s.count) {
    const bool vec = ((s.count | (int)(s.slab & 3)) & 3) == 0 && (reinterpret_cast<unsigned long long>(s.part) & 15ull) == 0;
    if (vec) {
#pragma unroll 5
      for (int k = w; k < s.n_slabs; k += 8) {
        const float4 x = *reinterpret_cast<const float4*>(s.part + (long long)k * s.slab + i);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
    } else {
      for (int k = w; k < s.n_slabs; k += 8) {
        const float* p = s.part + (long long)k * s.slab + i;
        acc.x += p[0];
        if (i + 1 < s.count) acc.y += p[1];
        if (i + 2 < s.count) acc.z += p[2];
        if (i + 3 < s.count) acc.w += p[3];
      }
    }
  }
  sh[w][lane] = acc;
  __syncthreads();
  if (w != 0 || i >= s.count) return;
  float4 t = sh[0][lane];
#pragma unroll
  for (int q = 1; q < 8; q++) { const float4 x = sh[q][lane]; t.x += x.x; t.y += x.y; t.z += x.z; t.w += x.w; }
  const float v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int ie = i + e;
    if (ie >= s.count) break;
    long long dst;
    if (s.kind == 0) {
      const int n = ie % s.N, mm = (ie / s.N) & 127, r = ie / (s.N * 128);
      const int ky = r / s.mts, mt = r - ky * s.mts;
      const int feat = mt * 128 + mm;
      if (feat >= s.KW * s.C) continue;
      int row = ky * s.KW * s.C + feat;
      if (s.s2d_k4) row = s2d_real_row(row, s.s2d_k4);
      dst = s.dst_off + (long long)row * s.N + n;
    } else dst = s.dst_off + ie;
    (s.dst_ptr ? s.dst_ptr : grads)[dst] = s.alpha * v[e];
  }
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
// split-K partial sums part[z][b][N] -> act(sum_z + bias): fp32 row-major and/or batch-planar planes.  FIN_ZL lanes share
// one 8-feature piece: lane zl sums slabs zl, zl + FIN_ZL, ... and a fixed xor tree combines them (bitwise reproducible).
constexpr int FIN_ZL = 8;
__global__ void bp_splitk_finish_kernel(const float* __restrict__ part, int n_z, long long part_z, int B, int N,
                                        const float* __restrict__ bias, int act, float* __restrict__ out_f32, BpT out) {
  pdl_wait(); pdl_trigger();
  const int b_pad = (B + 15) & ~15;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = t / FIN_ZL;
  const int zl = (int)(t - i * FIN_ZL);
  const int chunks = N >> 3;
  const int c = (int)(i / b_pad), b = (int)(i - (long long)c * b_pad);
  const bool valid = i < (long long)chunks * b_pad && b < B;
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (valid) {
    const float* p0 = part + (long long)b * N + c * 8;
#pragma unroll 4
    for (int z = zl; z < n_z; z += FIN_ZL) {
      const float4* p = reinterpret_cast<const float4*>(p0 + (long long)z * part_z);
      float4 x = p[0], y = p[1];
      v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w; v[4] += y.x; v[5] += y.y; v[6] += y.z; v[7] += y.w;
    }
  }
#pragma unroll
  for (int o = FIN_ZL / 2; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if (!valid || zl != 0) return;
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
// inverse of s2d_real_row: real HWIO row -> row of the space-to-depth ordered blob
__host__ __device__ inline int s2d_blob_row(int k, int k4) {
  const int c = k & 3, pix = k >> 2, x = pix % (4 * k4), y = pix / (4 * k4);
  return (((y >> 2) * k4 + (x >> 2)) << 6) + ((y & 3) << 4) + ((x & 3) << 2) + c;
}
// Blob planes of the parameters at flat index j .. j+cnt-1 (cnt <= 4, same weight row, n % 4 == 0 when cnt == 4):
// used by the optimiser kernel so that the blobs are refreshed by the same pass that updates the parameters.
__device__ __forceinline__ void blob_store1(const BlobSeg* __restrict__ segs, int n_segs, bf16* __restrict__ hi, long long lo_off,
                                            long long j, float p) {
  for (int q = 0; q < n_segs; q++) {
    const long long w0 = segs[q].w_off;
    const int K = segs[q].K, N = segs[q].N;
    if (j < w0 || j >= w0 + (long long)K * N) continue;
    const int rel = (int)(j - w0), k = rel / N, n = rel - k * N;
    const int kp = segs[q].s2d_k4 ? s2d_blob_row(k, segs[q].s2d_k4) : k;
    bf16* o = hi + segs[q].blob_off + ((long long)(n >> 3) * K + kp) * 8 + (n & 7);
    const bf16 h = __float2bfloat16_rn(p);
    o[0] = h; o[lo_off] = __float2bfloat16_rn(p - __bfloat162float(h));
    return;
  }
}
__device__ __forceinline__ void blob_store4(const BlobSeg* __restrict__ segs, int n_segs, bf16* __restrict__ hi, long long lo_off,
                                            long long j, const float p[4]) {
  for (int q = 0; q < n_segs; q++) {
    const long long w0 = segs[q].w_off, w1 = w0 + (long long)segs[q].K * segs[q].N;
    if (j + 3 < w0 || j >= w1) continue;
    const int K = segs[q].K, N = segs[q].N;
    const int rel = (int)(j - w0), k = rel / N, n = rel - k * N;
    if (j >= w0 && j + 3 < w1 && (n & 3) == 0) {          // the common case: four parameters of one weight row
      const int kp = segs[q].s2d_k4 ? s2d_blob_row(k, segs[q].s2d_k4) : k;
      bf16* o = hi + segs[q].blob_off + ((long long)(n >> 3) * K + kp) * 8 + (n & 7);
      uint32_t h0, l0, h1, l1;
      split2(p[0], p[1], h0, l0); split2(p[2], p[3], h1, l1);
      *reinterpret_cast<uint2*>(o) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(o + lo_off) = make_uint2(l0, l1);
    } else {
      for (int i = 0; i < 4; i++) blob_store1(segs, n_segs, hi, lo_off, j + i, p[i]);
    }
    return;
  }
}
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
// One block = one image row of 32 samples.  Phase 1 reads that row of every sample with 16-byte loads (contiguous
// 336-byte runs per sample) into shared memory; phase 2 hands every piece to one warp whose lanes are the 32 samples,
// so each store instruction writes 512 contiguous bytes of the plane.
constexpr int DEC_SAMPLES = 32;
constexpr int DEC_THREADS = 256;
__global__ void __launch_bounds__(DEC_THREADS)
bp_decode_s2d_kernel(const uint8_t* __restrict__ obs, const int32_t* __restrict__ idx, int B, int H, int W, int H4, int W4,
                     int padT, int padL, BpT dst) {
  // One block = one image row (canvas row Y, sub-row dy) of DEC_SAMPLES samples: 4*H4 x ceil(B/32) blocks, so that the
  // inference batch (B = 32: one sample group) still spreads over 84 blocks.
  extern __shared__ __align__(16) uint8_t dec_sm[];
  pdl_wait(); pdl_trigger();
  const int Y = blockIdx.x >> 2, dy = blockIdx.x & 3, b0 = blockIdx.y * DEC_SAMPLES;
  const int row_bytes = W * 4, units = row_bytes >> 4;          // 16-byte units per image row
  const int sstride = row_bytes + 8;                            // per-sample stride: +8 B keeps 8-byte lane reads conflict-free
  const int ns = min(DEC_SAMPLES, B - b0);
  const long long img = (long long)H * row_bytes;
  const int y = 4 * Y + dy - padT;
  const bool in_img = y >= 0 && y < H;
  for (int i = threadIdx.x; i < ns * units; i += DEC_THREADS) {
    const int sidx = i / units, ux = i - sidx * units;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (in_img) {
      const long long srow = idx ? idx[b0 + sidx] : b0 + sidx;
      v = *reinterpret_cast<const uint4*>(obs + srow * img + (long long)y * row_bytes + ux * 16);
    }
    uint2* d = reinterpret_cast<uint2*>(dec_sm + sidx * sstride + ux * 16);
    d[0] = make_uint2(v.x, v.y);
    d[1] = make_uint2(v.z, v.w);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half_units = W >> 1;                                // 8-byte units (pixel pairs) per image row
  for (int q = warp; q < W4 * 2; q += DEC_THREADS / 32) {
    const int X = q >> 1, dxh = q & 1;
    const int u8 = 2 * X + dxh - (padL >> 1);                   // pixel pair inside the source row
    uint4 o = make_uint4(0, 0, 0, 0);
    if (lane < ns && u8 >= 0 && u8 < half_units) {
      const uint2 v = *reinterpret_cast<const uint2*>(dec_sm + lane * sstride + u8 * 8);
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
