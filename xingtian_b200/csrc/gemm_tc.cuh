// gemm_tc.cuh -- tcgen05 (5th-gen tensor core) implicit-GEMM family for sm_100a.
//
//   D[128 x N] (fp32, TMEM) += A[128 x 16] * B[16 x N]   per tcgen05.mma.cta_group::1.kind::f16
//
// One warp-specialised kernel serves every conv / dense layer in forward, data-gradient and
// weight-gradient form:
//   warps 0-3 : producers.  Every tensor a GEMM consumes exists in HBM/L2 as two bf16 planes,
//               hi = bf16(x) and lo = bf16(x - hi) ("bf16x3": the tensor core computes
//               hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM, reproducing fp32 products to
//               ~2^-17, so the 1e-3 parity contract is met with margin while the math runs on
//               tcgen05).  The planes are written once by the kernel that produces the tensor
//               (GEMM epilogues, the uint8 frame decode, the optimiser for the weights), so a
//               producer thread only issues 16-byte cp.async copies (zero-filled where the im2col /
//               transposed-conv window leaves the image) into the UMMA canonical no-swizzle layout
//               and lets the copies arrive on the stage mbarrier asynchronously
//               (cp.async.mbarrier.arrive.noinc): no conversion work, no register staging, many
//               copies in flight per thread.  uint8 frames are exact in bf16 (no lo plane).
//   warp 4    : allocates TMEM, issues the MMAs (one elected lane), tcgen05.commit -> mbarriers.
//   warps 0-3 : epilogue after the K loop: tcgen05.ld (32 lanes x 16 columns per instruction),
//               bias+activation / activation-gradient mask / split-K atomic accumulation.
//
// Shared-memory tile layouts:
//   K-major operand  (rows = M/N entity, 64 K elements per stage): SWIZZLE_128B canonical layout -- row r is 128 B at
//       r*128, its 16-byte chunk c stored at position c ^ (r & 7); SBO (8-row group stride) = 1024 B; one K=16 step =
//       +32 B on the descriptor start address.  Lanes of a producer warp walk the 8 chunks of a row, so a warp
//       instruction reads 4 rows x 128 contiguous bytes and writes 4 conflict-free smem rows.
//   MN-major operand (rows = K entity, chunks along M/N): no-swizzle canonical layout addr(row,chunk) =
//       chunk*ROWS*16 + row*16; core matrix = 8 K-rows x 16 B contiguous, LBO (stride between 8-row K groups) = 128 B,
//       SBO (stride between 16-byte MN chunks) = ROWS*16.
// The im2col row gather therefore feeds the forward GEMM (K-major) and the weight-gradient GEMM (MN-major: the
// transposed use of the same rows) without any transpose.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_f32.cuh"

namespace xtb {
namespace tc {

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
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or ~hint ns pass)
// instead of burning issue slots -- the 5 consumer warps of a CTA would otherwise out-issue the producers
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
// 16-byte async copy global -> shared, zero-filled when src_bytes == 0
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival from this thread once all its prior cp.async have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (SWIZZLE_NONE, sm_100 version bits)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version = 1 (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// K-major operand in the SWIZZLE_128B canonical layout: rows of 128 B (64 bf16) at 128-byte stride, the 16-byte
// chunk index XORed with (row & 7); 8-row groups are 1024 B apart (SBO); the tile base is 1024-byte aligned and a
// K=16 step advances the start address by 32 B inside the swizzle atom.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major)
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32; // SBO
  d |= (uint64_t)1 << 46;                       // descriptor version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;                       // layout type SWIZZLE_128B
  return d;
}
// MN-major operand, 128 MN elements x 64 K rows, SWIZZLE_128B: K row r of MN-half h is 128 B at h*8192 + r*128 with
// its 16-byte chunk index XORed with (r & 7); LBO = 8192 (next 64 MN elements), SBO = 1024 (next 8 K rows); one
// K=16 step = two 8-row groups = +2048 B.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((8192u >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: bf16 x bf16 -> fp32, M=128
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// fp32 -> bf16 hi / residual lo
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
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t r[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
               "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_v8(const void* p, float r[8]) {
  uint32_t u[8];
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "l"(p));
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = __uint_as_float(u[i]);
}
// write 16 fp32 values (32-byte aligned run) with full-sector 256-bit stores
__device__ __forceinline__ void store_f32x16(float* dst, const float v[16]) {
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; i++) u[i] = __float_as_uint(v[i]);
  st_global_v8(dst, u);
#pragma unroll
  for (int i = 0; i < 8; i++) u[i] = __float_as_uint(v[8 + i]);
  st_global_v8(dst + 8, u);
}
// ... and their bf16 hi/lo planes (32 B each); hi_plane may be NULL (tensor has no tensor-core consumer)
__device__ __forceinline__ void store_planes16(__nv_bfloat16* hi_plane, long long lo_off, long long elem, const float v[16]) {
  if (!hi_plane) return;
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; i++) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  st_global_v8(hi_plane + elem, h);
  st_global_v8(hi_plane + lo_off + elem, l);
}
__device__ __forceinline__ void store_plane1(__nv_bfloat16* hi_plane, long long lo_off, long long elem, float x) {
  if (!hi_plane) return;
  __nv_bfloat16 h = __float2bfloat16_rn(x);
  hi_plane[elem] = h;
  hi_plane[lo_off + elem] = __float2bfloat16_rn(x - __bfloat162float(h));
}

// ------------------------------------------------------------------------------------------
// Chunk sources.  A source is a logical 2-D array [n_rows][n_cols] of bf16 whose rows are gathered
// independently and whose columns come in contiguous, 16-byte aligned groups of 8.
//   Row  : per-row state (init_row);  Col : per-column-group state (col_info holds the integer divisions;
//          the kernel caches it in shared memory once per tile when TABLE is set)
//   chunk(row, col) -> hi-plane address of the 8 elements, or NULL = zeros;  lo plane = hi + lo_off
// ------------------------------------------------------------------------------------------
constexpr int TC_PRODUCERS = 256;                 // 8 producer warps: the gathers are issue-latency bound per warp
constexpr int TC_MMA_WARP = TC_PRODUCERS / 32;    // warp 8
constexpr int TC_THREADS = TC_PRODUCERS + 32 + 128;   // + MMA warp + 4 epilogue warps (TMEM lane quarters 1,2,3,0)

typedef __nv_bfloat16 bf16;
struct Col { int a, b; };

// plain rows: X[row(m)*ld + col], optional row gather.  ld % 8 == 0.
// Row k' = ((ty*k4 + tx)*16 + dy*4 + dx)*4 + c of the weight matrix of the (k4 x k4, stride 1, 64 channel) conv over a
// space-to-depth plane is row ((4ty+dy)*4k4 + 4tx+dx)*4 + c of the real (4k4 x 4k4, stride 4, 4 channel) conv.
__device__ __forceinline__ int s2d_row(int m, int k4) {
  int tap = m >> 6, ty = tap / k4, tx = tap - ty * k4;
  int dy = (m >> 4) & 3, dx = (m >> 2) & 3, c = m & 3;
  return (((4 * ty + dy) * 4 * k4 + 4 * tx + dx) << 2) + c;
}
struct SrcDense {
  static constexpr bool TABLE = false;
  __device__ __forceinline__ void set_class(int) {}
  const bf16* x; long long lo_off; const int32_t* idx; int ld; int n_rows;
  int s2d_k4;     // > 0: row m is a space-to-depth weight row (see s2d_row), read the real row instead
  struct Row { long long base; int valid; };
  __device__ __forceinline__ const bf16* any() const { return x; }
  __device__ __forceinline__ void init_row(int m, Row& r) const {
    r.valid = m < n_rows;
    int src = s2d_k4 ? s2d_row(m, s2d_k4) : m;
    r.base = r.valid ? (long long)(idx ? idx[src] : src) * ld : 0;
  }
  __device__ __forceinline__ Col col_info(int col) const { return Col{col, 0}; }
  __device__ __forceinline__ const bf16* chunk(const Row& r, const Col& c) const { return r.valid ? x + r.base + c.a : nullptr; }
};

// im2col rows: row = output position (b,oy,ox), col = (ky,kx,ci).  C is a power of two (cshift = log2 C).
// C >= 8: a chunk is 8 channels of one tap;  C == 4: a chunk is 2 horizontally adjacent taps (KW, padL and the
// stride are even so both taps are inside or both outside the image, and the 8 values are contiguous).
template <bool PAD>
struct SrcIm2col {
  static constexpr bool TABLE = true;
  __device__ __forceinline__ void set_class(int) {}
  const bf16* x; long long lo_off; ConvGeom g; int cshift; int n_rows;
  struct Row { long long base; int iy0, ix0, valid; };
  __device__ __forceinline__ const bf16* any() const { return x; }
  __device__ __forceinline__ void init_row(int m, Row& r) const {
    r.valid = m < n_rows;
    if (!r.valid) { r.base = 0; r.iy0 = r.ix0 = 0; return; }
    int b = fastdiv(m, g.mP); int p = m - b * g.P; int oy = fastdiv(p, g.mOW); int ox = p - oy * g.OW;
    r.iy0 = oy * g.S - g.padT; r.ix0 = ox * g.S - g.padL;
    r.base = (((long long)b * g.H + r.iy0) * g.W + r.ix0) * g.C;
  }
  __device__ __forceinline__ Col col_info(int col) const {
    int tap = col >> cshift, ci = col & (g.C - 1);
    int ky = tap / g.KW, kx = tap - ky * g.KW;
    return Col{((ky * g.W + kx) << cshift) + ci, (ky << 16) | kx};
  }
  __device__ __forceinline__ const bf16* chunk(const Row& r, const Col& c) const {
    if (!r.valid) return nullptr;
    if (PAD) {
      int iy = r.iy0 + (c.b >> 16), ix = r.ix0 + (c.b & 0xffff);
      if ((unsigned)iy >= (unsigned)g.H || (unsigned)ix >= (unsigned)g.W) return nullptr;
    }
    return x + r.base + c.a;
  }
};

// transposed-conv gather: row = input position (b,iy,ix), col = (ky,kx,co); value = dY[b,oy,ox,co]
// with oy=(iy+padT-ky)/S when divisible and in range.  Cout is a power of two >= 8 (oshift = log2 Cout).
struct SrcDgrad {
  static constexpr bool TABLE = true;
  __device__ __forceinline__ void set_class(int) {}
  const bf16* dy; long long lo_off; ConvGeom g; int oshift; int sshift; int n_rows;
  struct Row { long long base; int iy0, ix0, valid; };
  __device__ __forceinline__ const bf16* any() const { return dy; }
  __device__ __forceinline__ void init_row(int m, Row& r) const {
    r.valid = m < n_rows;
    if (!r.valid) { r.base = 0; r.iy0 = r.ix0 = 0; return; }
    int HW = g.H * g.W; int b = fastdiv(m, g.mHW); int p = m - b * HW; int iy = fastdiv(p, g.mW); int ix = p - iy * g.W;
    r.base = (long long)b * g.P * g.Cout; r.iy0 = iy + g.padT; r.ix0 = ix + g.padL;
  }
  __device__ __forceinline__ Col col_info(int col) const {
    int tap = col >> oshift, co = col & (g.Cout - 1);
    int ky = tap / g.KW, kx = tap - ky * g.KW;
    return Col{co, (ky << 16) | kx};
  }
  __device__ __forceinline__ const bf16* chunk(const Row& r, const Col& c) const {
    if (!r.valid) return nullptr;
    int ty = r.iy0 - (c.b >> 16), tx = r.ix0 - (c.b & 0xffff);
    int mask = g.S - 1;
    if (ty < 0 || tx < 0 || (ty & mask) || (tx & mask)) return nullptr;
    int oy = ty >> sshift, ox = tx >> sshift;
    if (oy >= g.OH || ox >= g.OW) return nullptr;
    return dy + r.base + (((long long)(oy * g.OW + ox)) << oshift) + c.a;
  }
};

// conv weights for the data gradient: row = ci, col = (ky,kx,co) -> W[ky,kx,ci,co]
struct SrcWDgrad {
  static constexpr bool TABLE = false;
  __device__ __forceinline__ void set_class(int) {}
  const bf16* w; long long lo_off; int Cin, Cout, oshift; int n_rows;
  struct Row { int ci, valid; };
  __device__ __forceinline__ const bf16* any() const { return w; }
  __device__ __forceinline__ void init_row(int m, Row& r) const { r.valid = m < n_rows; r.ci = m; }
  __device__ __forceinline__ Col col_info(int col) const { return Col{col >> oshift, col & (Cout - 1)}; }
  __device__ __forceinline__ const bf16* chunk(const Row& r, const Col& c) const {
    if (!r.valid) return nullptr;
    return w + ((long long)(c.a * Cin + r.ci) << oshift) + c.b;
  }
};

// Stride-S transposed-conv gather decomposed by parity class (py,px) of (iy+padT, ix+padL): only the taps with
// ky = py (mod S), kx = px (mod S) reach an input position of that class, so each class is a dense GEMM with
// K/S^2 columns over 1/S^2 of the rows instead of a GEMM whose operand tile is (S^2-1)/S^2 zeros.
// Rows of class z are enumerated (b, jy, jx): uy = uy0 + S*jy, ux = ux0 + S*jx;  cols = (ty, tx, co) with
// ky = py + S*ty, kx = px + S*tx.  KH % S == 0, KW % S == 0.
struct DgradClass {
  unsigned long long mPer, mNX;                 // fastdiv magics of nY*nX, nX
  int S, sshift, padT, padL, H, W, nY, nX;      // nY = ceil(H/S), nX = ceil(W/S)
  int py, px, uy0, ux0;
  __device__ __forceinline__ void set(int z) {
    py = z >> sshift; px = z & (S - 1);
    uy0 = padT + ((py - padT) & (S - 1));
    ux0 = padL + ((px - padL) & (S - 1));
  }
  // class row m -> (b, iy, ix); returns false when the position falls outside the image (ragged classes)
  __device__ __forceinline__ bool decode(int m, int& b, int& iy, int& ix, int& qy, int& qx) const {
    int per = nY * nX;
    b = fastdiv(m, mPer); int p = m - b * per; int jy = fastdiv(p, mNX), jx = p - jy * nX;
    int uy = uy0 + S * jy, ux = ux0 + S * jx;
    iy = uy - padT; ix = ux - padL;
    qy = (uy - py) >> sshift; qx = (ux - px) >> sshift;
    return iy < H && ix < W;
  }
};
struct SrcDgradS {
  static constexpr bool TABLE = true;
  const bf16* dy; long long lo_off; ConvGeom g; int oshift; int n_rows; DgradClass k; int KWs;
  struct Row { long long base; int qy, qx, valid; };
  __device__ __forceinline__ const bf16* any() const { return dy; }
  __device__ __forceinline__ void set_class(int z) { k.set(z); }
  __device__ __forceinline__ void init_row(int m, Row& r) const {
    r.valid = m < n_rows; r.base = 0; r.qy = r.qx = 0;
    if (!r.valid) return;
    int b, iy, ix;
    r.valid = k.decode(m, b, iy, ix, r.qy, r.qx);
    r.base = (long long)b * g.P * g.Cout;
  }
  __device__ __forceinline__ Col col_info(int col) const {
    int tap = col >> oshift, co = col & (g.Cout - 1);
    int ty = tap / KWs, tx = tap - ty * KWs;
    return Col{co, (ty << 16) | tx};
  }
  __device__ __forceinline__ const bf16* chunk(const Row& r, const Col& c) const {
    if (!r.valid) return nullptr;
    int oy = r.qy - (c.b >> 16), ox = r.qx - (c.b & 0xffff);
    if ((unsigned)oy >= (unsigned)g.OH || (unsigned)ox >= (unsigned)g.OW) return nullptr;
    return dy + r.base + (((long long)(oy * g.OW + ox)) << oshift) + c.a;
  }
};
struct SrcWDgradS {
  static constexpr bool TABLE = true;
  const bf16* w; long long lo_off; int Cin, Cout, oshift; int n_rows; int KW, KWs; DgradClass k;
  struct Row { int ci, valid; };
  __device__ __forceinline__ const bf16* any() const { return w; }
  __device__ __forceinline__ void set_class(int z) { k.set(z); }
  __device__ __forceinline__ void init_row(int m, Row& r) const { r.valid = m < n_rows; r.ci = m; }
  __device__ __forceinline__ Col col_info(int col) const {
    int tap = col >> oshift, co = col & (Cout - 1);
    int ty = tap / KWs, tx = tap - ty * KWs;
    int ky = k.py + k.S * ty, kx = k.px + k.S * tx;
    return Col{(ky * KW + kx) * Cin, co};
  }
  __device__ __forceinline__ const bf16* chunk(const Row& r, const Col& c) const {
    if (!r.valid) return nullptr;
    return w + ((long long)(c.a + r.ci) << oshift) + c.b;
  }
};

// ------------------------------------------------------------------------------------------
// Epilogues: store16(row m, first column n, 16 accumulators).  Besides the fp32 tensor they keep the
// bf16 hi/lo planes of the result current (hi == NULL when nothing on the tensor-core path reads it).
// Rows are written as 32-byte full-sector stores (ld % 16 == 0 on the fast path).
// ------------------------------------------------------------------------------------------
struct EpiNoState {};
struct TcEpiBiasAct {
  typedef EpiNoState State;
  float* out; const float* bias; float alpha; int act; int ld; int M, N; bf16* hi; long long lo_off;
  __device__ __forceinline__ void set_class(int) {}
  __device__ __forceinline__ void init(State&) const {}
  __device__ __forceinline__ void finish(State&, float*, int) const {}
  __device__ __forceinline__ void store16(State&, int, int m, int n0, const float v[16]) const {
    if (m >= M || n0 >= N) return;
    long long e = (long long)m * ld + n0;
    float r[16];
#pragma unroll
    for (int j = 0; j < 16; j++) r[j] = (n0 + j < N) ? act_apply(act, alpha * v[j] + bias[n0 + j]) : 0.f;
    if (n0 + 15 < N && (ld & 15) == 0) {
      store_f32x16(out + e, r);
      store_planes16(hi, lo_off, e, r);
    } else {
      for (int j = 0; j < 16; j++) if (n0 + j < N) { out[e + j] = r[j]; store_plane1(hi, lo_off, e + j, r[j]); }
    }
  }
};
// `db` (optional): bias gradient of the layer that produced the source tensor = column sums of this gradient.
// Each epilogue thread keeps per-column partial sums in registers across all tiles of its persistent CTA; they
// are reduced once (warp shuffles -> shared memory -> one atomic per column per CTA).  Requires a single N tile.
struct TcEpiDgrad {
  struct State { float acc[64]; };
  float* gout; const float* srcout; int act; int ld; int accumulate; int M, N; bf16* hi; long long lo_off; float* db;
  int use_class; DgradClass k;     // use_class: rows are parity-class rows (b,jy,jx) of a strided transposed conv
  __device__ __forceinline__ void set_class(int z) { if (use_class) k.set(z); }
  __device__ __forceinline__ void init(State& st) const {
    if (db) {
#pragma unroll
      for (int j = 0; j < 64; j++) st.acc[j] = 0.f;
    }
  }
  __device__ __forceinline__ void store16(State& st, int c, int m, int n0, const float v[16]) const {
    if (m >= M || n0 >= N) return;
    long long e = (long long)m * ld + n0;
    if (use_class) {
      int b, iy, ix, qy, qx;
      if (!k.decode(m, b, iy, ix, qy, qx)) return;
      e = (((long long)b * k.H + iy) * k.W + ix) * ld + n0;
    }
    float r[16];
    if (n0 + 15 < N && (ld & 15) == 0) {
      float so[16], go[16];
      ld_global_v8(srcout + e, so); ld_global_v8(srcout + e + 8, so + 8);
      if (accumulate) { ld_global_v8(gout + e, go); ld_global_v8(gout + e + 8, go + 8); }
#pragma unroll
      for (int j = 0; j < 16; j++) {
        float gq = v[j] * act_grad_from_out(act, so[j]);
        r[j] = accumulate ? go[j] + gq : gq;
        if (db) {   // c is a compile-time multiple of 16 after unrolling of the caller's column loop
#pragma unroll
          for (int q = 0; q < 4; q++) if (c == 16 * q) st.acc[16 * q + j] += gq;
        }
      }
      store_f32x16(gout + e, r);
      store_planes16(hi, lo_off, e, r);
    } else {
      for (int j = 0; j < 16; j++) if (n0 + j < N) {
        float gq = v[j] * act_grad_from_out(act, srcout[e + j]);
        float rr = accumulate ? gout[e + j] + gq : gq;
        gout[e + j] = rr; store_plane1(hi, lo_off, e + j, rr);
      }
    }
  }
  // sh: zero-initialised shared floats [>= N]; called once by all 128 epilogue threads
  __device__ __forceinline__ void finish(State& st, float* sh, int nt) const {
    if (!db) return;
#pragma unroll
    for (int j = 0; j < 64; j++) {
      if (j < nt) {
        float q = st.acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(sh + j, q);
      }
    }
    asm volatile("bar.sync 2, 128;" ::: "memory");
    int t = (int)threadIdx.x - (TC_PRODUCERS + 32);
    if (t >= 0 && t < nt && t < N) atomicAdd(db + t, sh[t]);
  }
};
struct TcEpiStore {     // plain row-major store (self-test)
  typedef EpiNoState State;
  float* out; int ld; int M, N;
  __device__ __forceinline__ void set_class(int) {}
  __device__ __forceinline__ void init(State&) const {}
  __device__ __forceinline__ void finish(State&, float*, int) const {}
  __device__ __forceinline__ void store16(State&, int, int m, int n0, const float v[16]) const {
    if (m >= M) return;
#pragma unroll
    for (int j = 0; j < 16; j++)
      if (n0 + j < N) out[(long long)m * ld + n0 + j] = v[j];
  }
};
struct TcEpiAtomic {
  typedef EpiNoState State;
  float* out; float alpha; int ld; int M, N;
  int s2d_k4;                  // > 0: GEMM row m accumulates into output row s2d_row(m) (space-to-depth weight order)
  __device__ __forceinline__ void set_class(int) {}
  __device__ __forceinline__ void init(State&) const {}
  __device__ __forceinline__ void finish(State&, float*, int) const {}
  __device__ __forceinline__ void store16(State&, int, int m, int n0, const float v[16]) const {
    if (m >= M) return;
    float* o = out + (long long)(s2d_k4 ? s2d_row(m, s2d_k4) : m) * ld + n0;
#pragma unroll
    for (int j = 0; j < 16; j++)
      if (n0 + j < N) atomicAdd(o + j, alpha * v[j]);
  }
};

// ------------------------------------------------------------------------------------------
// The kernel: persistent, warp-specialised.
//   warps 0-7  producers   (cp.async gathers into the stage ring)
//   warp  8    MMA issuer  (tcgen05.mma into one of two TMEM accumulator buffers)
//   warps 9-12 epilogue    (tcgen05.ld of the other buffer -> global) -- overlaps the next tile's main loop
// Each CTA walks tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...  of the (m-tile, n-tile, k-split) space.
//   A_MN / B_MN : operand is MN-major (its tile rows are the K entity) instead of K-major.
//   KT          : reduction elements per pipeline stage;  NT : N tile = UMMA N (16..256, multiple of 16)
//   A_SPLIT / B_SPLIT : operand carries a bf16 residual plane (bf16x3); false = exact in bf16
// ------------------------------------------------------------------------------------------

// Optional timeline trace (XTB_TC_TRACE builds only): %globaltimer stamps per CTA / role / tile event.
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#ifdef XTB_TC_TRACE
__device__ unsigned long long* g_tc_trace = nullptr;     // [cta][slot] ; slot 0 = count
constexpr int TC_TRACE_SLOTS = 256;
__device__ __forceinline__ void trace(int code) {
  if (!g_tc_trace) return;
  unsigned long long* b = g_tc_trace + (size_t)blockIdx.x * TC_TRACE_SLOTS;
  unsigned long long i = atomicAdd(b, 1ULL) + 1;
  if (i < TC_TRACE_SLOTS) b[i] = (gtimer() << 8) | (unsigned)code;
}
#define TC_TRACE(code) do { if ((threadIdx.x & 31) == 0) trace(code); } while (0)
#else
#define TC_TRACE(code) do {} while (0)
#endif
// Optional per-role cycle accounting (XTB_TC_PROF builds only): clock64 deltas accumulated in registers by the lead
// warp of each role and written once at kernel exit ([cta][role*8 + slot]); no memory traffic inside the loops.
#ifdef XTB_TC_PROF
__device__ long long* g_tc_prof = nullptr;
#define PROF_DECL long long pf_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pf_last = clock64();
#define PROF(i) do { long long n_ = clock64(); pf_acc[i] += n_ - pf_last; pf_last = n_; } while (0)
#define PROF_STAMP(slot) do { if (g_tc_prof && threadIdx.x == 0) g_tc_prof[(size_t)blockIdx.x * 32 + (slot)] = (long long)gtimer(); } while (0)
#define PROF_FLUSH(role) do { if (g_tc_prof && (threadIdx.x & 31) == 0) { for (int q_ = 0; q_ < 8; q_++) g_tc_prof[(size_t)blockIdx.x * 32 + (role) * 8 + q_] = pf_acc[q_]; } } while (0)
#else
#define PROF_DECL
#define PROF(i) do {} while (0)
#define PROF_FLUSH(role) do {} while (0)
#define PROF_STAMP(slot) do {} while (0)
#endif
constexpr int TC_MAXCOLS = 128;   // cached column groups per operand (K-major conv operands: K/8 <= 128)

template <int NT, int KT, bool A_MN, bool B_MN, bool A_SPLIT, bool B_SPLIT>
struct TcCfg {
  static constexpr int A_ROWS = A_MN ? KT : 128;
  static constexpr int A_CHUNKS = A_MN ? 16 : KT / 8;
  static constexpr int B_ROWS = B_MN ? KT : NT;
  static constexpr int B_CHUNKS = B_MN ? NT / 8 : KT / 8;
  static constexpr int A_BYTES = A_ROWS * A_CHUNKS * 16;
  static constexpr int B_BYTES = B_ROWS * B_CHUNKS * 16;
  static constexpr int STAGE_BYTES = A_BYTES * (A_SPLIT ? 2 : 1) + B_BYTES * (B_SPLIT ? 2 : 1);
  static constexpr int STAGES = (STAGE_BYTES * 4 <= 108 * 1024) ? 4 : ((STAGE_BYTES * 3 <= 108 * 1024) ? 3 : 2);
  static constexpr int SMEM_BYTES = STAGE_BYTES * STAGES + 1024;
  // B_SPLIT: one MMA of N = 2*NT multiplies A by [B_hi | B_lo] (the planes are adjacent N chunks in shared memory), so an
  // accumulator buffer is 2*NT columns and the epilogue adds the halves; A tiles are read once instead of twice
  static constexpr int ACC_COLS = B_SPLIT ? 2 * NT : NT;
  static_assert(ACC_COLS <= 256, "accumulator tile too wide");
  static constexpr int TMEM_COLS = ACC_COLS <= 16 ? 32 : (ACC_COLS <= 32 ? 64 : (ACC_COLS <= 64 ? 128 : (ACC_COLS <= 128 ? 256 : 512)));  // two buffers
  static constexpr int CTAS_PER_SM = (2 * SMEM_BYTES + 8192 <= 227 * 1024 && TMEM_COLS <= 256) ? 2 : 1;
};

// Per-operand producer.
//  MN-major operand (tile rows = K entity): no-swizzle layout addr(row,chunk) = chunk*ROWS*16 + row*16, lanes run
//    along rows (consecutive K rows are adjacent in memory for weights / dY), row state in registers per stage.
//  K-major operand (tile rows = M/N entity, KT = 64): SWIZZLE_128B layout, lanes run along the 8 chunks of a row so
//    that a warp instruction reads 4 rows x 128 contiguous bytes (instead of 32 scattered 16-byte pieces, which made
//    the gather L1-tag bound) and writes 4 conflict-free 128-byte rows; row state is held in registers for the whole output tile.
template <int ROWS, int CHUNKS, bool MN, bool SPLIT, class SRC>
struct Producer {
  static constexpr int SLOTS = ROWS * CHUNKS;
  static_assert(MN || CHUNKS == 8, "K-major tiles are 64 elements (one 128-byte swizzle atom) wide");
  static constexpr int PER = (SLOTS + TC_PRODUCERS - 1) / TC_PRODUCERS;
  static constexpr int NROW = MN ? PER : 1;          // distinct K rows per thread per stage (MN-major, no swizzle)
  static constexpr int RPPM = TC_PRODUCERS / CHUNKS;   // rows per pass
  static_assert(!MN || CHUNKS == 16 || TC_PRODUCERS % CHUNKS == 0, "a thread keeps one chunk column");
  typename SRC::Row rows[NROW];

  // ---- MN-major, no swizzle (weights / dY rows).  Lanes run along the chunks of a K row first, so a warp reads
  // whole 128-byte lines (every CTA of the grid reads the same weight tile at the same time: with lanes along the
  // rows each instruction asked L2 for 32 separate sectors of 16 useful bytes and the shared lines became an L2
  // hot spot that paced the whole pipeline).
  __device__ __forceinline__ void init_rows(const SRC& src, int row0, int row_lim, int tid) {
#pragma unroll
    for (int q = 0; q < NROW; q++) {
      int r = tid / CHUNKS + q * RPPM;
      src.init_row(row0 + r, rows[q]);
      if (row0 + r >= row_lim || r >= ROWS) rows[q].valid = 0;
    }
  }
  __device__ __forceinline__ void produce(const SRC& src, const Col* cols, int col_base, uint32_t hi_smem, uint32_t lo_smem,
                                          int col0, int col_lim, int tid) {
    const bf16* dummy = src.any();
    const int c = tid % CHUNKS;
    const int gcol = col0 + 8 * c;
    const bool col_ok = gcol < col_lim;
    Col ci = Col{0, 0};
    if (col_ok) ci = SRC::TABLE ? cols[(gcol - col_base) >> 3] : src.col_info(gcol);
#pragma unroll
    for (int i = 0; i < NROW; i++) {
      int r = tid / CHUNKS + i * RPPM;
      if (ROWS % RPPM != 0 && r >= ROWS) break;
      const bf16* p = col_ok ? src.chunk(rows[i], ci) : nullptr;
      uint32_t off = (uint32_t)c * ROWS * 16 + (uint32_t)r * 16;
      cp_async16(hi_smem + off, p ? p : dummy, p ? 16u : 0u);
      if (SPLIT) cp_async16(lo_smem + off, p ? p + src.lo_off : dummy, p ? 16u : 0u);
    }
  }
  // ---- MN-major, 128 MN elements (the A operand of a weight gradient): swizzled, lanes along the 16 chunks of a row
  static __device__ __forceinline__ void produce_mn16(const SRC& src, const typename SRC::Row* rows_sm, const Col* cols, int col_base,
                                                      uint32_t hi_smem, uint32_t lo_smem, int mn0, int mn_ext, int tid) {
    const bf16* dummy = src.any();
    const int c = tid & 15;
    const int gcol = mn0 + 8 * c;
    Col ci = Col{0, 0};
    const bool col_ok = gcol < mn_ext;
    if (col_ok) ci = SRC::TABLE ? cols[(gcol - col_base) >> 3] : src.col_info(gcol);
    const uint32_t half = (uint32_t)(c >> 3) * 8192u;
    constexpr int RPP = TC_PRODUCERS / 16;      // rows per pass
#pragma unroll
    for (int i = 0; i < (ROWS + RPP - 1) / RPP; i++) {
      int r = (tid >> 4) + i * RPP;
      if (ROWS % RPP != 0 && r >= ROWS) break;
      const bf16* p = col_ok ? src.chunk(rows_sm[r], ci) : nullptr;
      uint32_t off = half + (uint32_t)r * 128 + (uint32_t)(((c & 7) ^ (r & 7)) << 4);
      cp_async16(hi_smem + off, p ? p : dummy, p ? 16u : 0u);
      if (SPLIT) cp_async16(lo_smem + off, p ? p + src.lo_off : dummy, p ? 16u : 0u);
    }
  }
  // ---- shared-memory row table (MN-major A of a weight gradient: the K rows change every stage)
  static __device__ __forceinline__ void fill_rows(const SRC& src, typename SRC::Row* rows_sm, int row0, int row_lim, int tid) {
    for (int r = tid; r < ROWS; r += TC_PRODUCERS) {
      typename SRC::Row rw;
      src.init_row(row0 + r, rw);
      if (row0 + r >= row_lim) rw.valid = 0;
      rows_sm[r] = rw;
    }
  }
  // ---- K-major: swizzled, lanes along the 8 chunks of a row.  A thread serves the same rows (tid/8 + i*RPPK) for
  // every K tile of an output tile, so their state lives in registers for the whole tile: the per-stage work is
  // one column lookup plus address arithmetic (row state in shared memory cost two dependent LDS round trips per
  // 16-byte copy, which is what the producers stalled on).
  static constexpr int RPPK = TC_PRODUCERS / 8;       // rows per pass
  static constexpr int KROWS = MN ? 1 : (ROWS + RPPK - 1) / RPPK;
  typename SRC::Row krows[KROWS];
  __device__ __forceinline__ void load_rows_k(const SRC& src, int row0, int row_lim, int tid) {
#pragma unroll
    for (int i = 0; i < KROWS; i++) {
      int r = (tid >> 3) + i * RPPK;
      src.init_row(row0 + r, krows[i]);
      if (row0 + r >= row_lim || r >= ROWS) krows[i].valid = 0;
    }
  }
  __device__ __forceinline__ void produce_k(const SRC& src, const Col* cols, int col_base, uint32_t hi_smem, uint32_t lo_smem,
                                            int k0, int k_end, int tid) const {
    const bf16* dummy = src.any();
    const int c = tid & 7;
    const int gcol = k0 + 8 * c;
    Col ci = Col{0, 0};
    const bool col_ok = gcol < k_end;
    if (col_ok) ci = SRC::TABLE ? cols[(gcol - col_base) >> 3] : src.col_info(gcol);
#pragma unroll
    for (int i = 0; i < KROWS; i++) {
      int r = (tid >> 3) + i * RPPK;
      if (ROWS % RPPK != 0 && r >= ROWS) break;
      const bf16* p = col_ok ? src.chunk(krows[i], ci) : nullptr;
      uint32_t off = (uint32_t)r * 128 + (uint32_t)((c ^ (r & 7)) << 4);
      cp_async16(hi_smem + off, p ? p : dummy, p ? 16u : 0u);
      if (SPLIT) cp_async16(lo_smem + off, p ? p + src.lo_off : dummy, p ? 16u : 0u);
    }
  }
};

__device__ __forceinline__ void producer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_PRODUCERS) : "memory"); }

// CLASSES: the third tile coordinate selects a parity class (set_class) instead of a K split
template <int NT, int KT, bool A_MN, bool B_MN, bool A_SPLIT, bool B_SPLIT, bool CLASSES, class ASRC, class BSRC, class EPI>
__global__ void __launch_bounds__(TC_THREADS)
tc_gemm_kernel(ASRC asrc, BSRC bsrc, EPI epi, int m_ext, int n_ext, int k_ext, int k_chunk, int m_tiles, int n_tiles,
               int total_tiles) {
  using Cfg = TcCfg<NT, KT, A_MN, B_MN, A_SPLIT, B_SPLIT>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ Col col_tab[2][2][TC_MAXCOLS];     // [tile parity][operand][column group]
  __shared__ float epi_sh[64];                  // epilogue reductions (bias gradients)
  __shared__ typename ASRC::Row a_rows[2][A_MN ? KT : 1];   // MN-major A: per-stage row state (K-major rows live in registers)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  PROF_STAMP(24);     // CTA entry
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * STAGES]), tempty0 = smem_u32(&bars[2 * STAGES + 2]);
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, TC_PRODUCERS); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; b++) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 128); }
    fence_barrier_init();
  }
  if (warp == TC_MMA_WARP) tmem_alloc(smem_u32(&tmem_slot), Cfg::TMEM_COLS);
  if (tid < 64) epi_sh[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // prologue done (barriers, TMEM, no global memory touched): wait for the producer grid, then let the next
  // kernel start its own prologue underneath this one
  TC_TRACE(1);      // prologue done
  PROF_STAMP(25);
  pdl_wait();
  pdl_trigger();
  PROF_STAMP(26);
  TC_TRACE(2);      // dependency satisfied
  const uint32_t tmem_base = tmem_slot;
  const uint32_t smem_base = smem_u32(smem);
  const int mn_tiles = m_tiles * n_tiles;

  if (warp < TC_MMA_WARP) {
    // ================= producers =================
    Producer<Cfg::A_ROWS, Cfg::A_CHUNKS, A_MN, A_SPLIT, ASRC> pa;
    Producer<Cfg::B_ROWS, Cfg::B_CHUNKS, B_MN, B_SPLIT, BSRC> pb;
    int stage = 0; uint32_t phase = 0; int it = 0;
    PROF_DECL
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
      const int z = tile / mn_tiles, rem = tile - z * mn_tiles;
      const int nt_i = rem / m_tiles, mt_i = rem - nt_i * m_tiles;
      const int m0 = mt_i * 128, n0 = nt_i * NT;
      const int kbeg = CLASSES ? 0 : z * k_chunk, kend = CLASSES ? k_ext : min(k_ext, kbeg + k_chunk);
      const int n_kt = (kend - kbeg + KT - 1) / KT;
      if (CLASSES) { asrc.set_class(z); bsrc.set_class(z); }
      Col* ta = col_tab[it & 1][0];
      Col* tb = col_tab[it & 1][1];
      // column-group tables: K-major operands cache the K range of this tile, MN-major ones their MN range
      if (ASRC::TABLE) {
        int base = A_MN ? m0 : kbeg, lim = A_MN ? min(m_ext, m0 + 128) : kend;
        for (int j = tid; base + 8 * j < lim && j < TC_MAXCOLS; j += TC_PRODUCERS) ta[j] = asrc.col_info(base + 8 * j);
      }
      if (BSRC::TABLE) {
        int base = B_MN ? n0 : kbeg, lim = B_MN ? min(n_ext, n0 + NT) : kend;
        for (int j = tid; base + 8 * j < lim && j < TC_MAXCOLS; j += TC_PRODUCERS) tb[j] = bsrc.col_info(base + 8 * j);
      }
      if (!A_MN) pa.load_rows_k(asrc, m0, m_ext, tid);
      if (!B_MN) pb.load_rows_k(bsrc, n0, n_ext, tid);
      if (ASRC::TABLE || BSRC::TABLE) producer_bar_sync();
      PROF(0);
      for (int t = 0; t < n_kt; t++) {
        const int k0 = kbeg + t * KT;
        if (A_MN) { pa.fill_rows(asrc, a_rows[t & 1], k0, kend, tid); producer_bar_sync(); }
        if (warp == 0) TC_TRACE(11);
        if (B_MN) pb.init_rows(bsrc, k0, kend, tid);
        PROF(4);
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        PROF(1);
        if (warp == 0) TC_TRACE(12);
        uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
        uint32_t sa_lo = sa + Cfg::A_BYTES;
        uint32_t sb = sa + Cfg::A_BYTES * (A_SPLIT ? 2 : 1);
        uint32_t sb_lo = sb + Cfg::B_BYTES;
#ifndef XTB_EXP_NO_A
        if (A_MN) pa.produce_mn16(asrc, a_rows[t & 1], ta, m0, sa, sa_lo, m0, m_ext, tid);
        else pa.produce_k(asrc, ta, kbeg, sa, sa_lo, k0, kend, tid);
#endif
        if (warp == 0) TC_TRACE(13);
        PROF(2);
#ifndef XTB_EXP_NO_B
        if (B_MN) pb.produce(bsrc, tb, n0, sb, sb_lo, n0, n_ext, tid);
        else pb.produce_k(bsrc, tb, kbeg, sb, sb_lo, k0, kend, tid);
#endif
        PROF(3);
        cp_async_mbar_arrive_noinc(full0 + 8 * stage);
        PROF(5);
        if (warp == 0) TC_TRACE(10);   // stage issued
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (warp == 0) PROF_FLUSH(0);
  } else if (warp == TC_MMA_WARP) {
    // ================= MMA issuer =================
    constexpr uint32_t IDESC = make_idesc(NT, A_MN, B_MN);
    constexpr uint32_t IDESC_HL = make_idesc(Cfg::ACC_COLS, A_MN, B_MN);    // N covers [B_hi | B_lo]
    constexpr uint32_t A_LBO = A_MN ? 128 : Cfg::A_ROWS * 16, A_SBO = A_MN ? Cfg::A_ROWS * 16 : 128;
    constexpr uint32_t B_LBO = B_MN ? 128 : Cfg::B_ROWS * 16, B_SBO = B_MN ? Cfg::B_ROWS * 16 : 128;
    constexpr uint32_t A_KSTEP = A_MN ? 256 : 2 * Cfg::A_ROWS * 16;   // advance of one K=16 step
    constexpr uint32_t B_KSTEP = B_MN ? 256 : 2 * Cfg::B_ROWS * 16;
    int stage = 0; uint32_t phase = 0; int it = 0;
    PROF_DECL
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
      const int z = tile / mn_tiles;
      const int kbeg = CLASSES ? 0 : z * k_chunk, kend = CLASSES ? k_ext : min(k_ext, kbeg + k_chunk);
      const int n_kt = (kend - kbeg + KT - 1) / KT;
      const int buf = it & 1;
      const uint32_t acc = tmem_base + buf * Cfg::ACC_COLS;
      PROF(3);
      mbar_wait(tempty0 + 8 * buf, ((it >> 1) & 1) ^ 1);      // the epilogue drained this accumulator
      tc_fence_after();
      PROF(0);
      for (int t = 0; t < n_kt; t++) {
        mbar_wait(full0 + 8 * stage, phase);
        PROF(1);
        TC_TRACE(20);                  // stage data landed
        if (lane == 0) {
#ifndef XTB_EXP_NO_PFENCE
          fence_proxy_async();      // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
#endif
          tc_fence_after();
          uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          uint32_t sa_lo = sa + Cfg::A_BYTES;
          uint32_t sb = sa + Cfg::A_BYTES * (A_SPLIT ? 2 : 1);
          uint32_t sb_lo = sb + Cfg::B_BYTES;
#ifndef XTB_EXP_NO_MMA
#pragma unroll
          for (int j = 0; j < KT / 16; j++) {
            // MN-major: no-swizzle canonical tile, 16 K rows per step; K-major: 128B-swizzled tile, 32 B per step
            uint64_t ah = A_MN ? make_desc_mn_sw128(sa + j * 2048) : make_desc_sw128(sa + j * 32);
            uint64_t bh = B_MN ? make_desc(sb + j * B_KSTEP, B_LBO, B_SBO) : make_desc_sw128(sb + j * 32);
            // columns [0,NT) += A_hi*B_hi, columns [NT,2NT) += A_hi*B_lo in one instruction (the lo tile continues the hi tile)
            umma_bf16(acc, ah, bh, IDESC_HL, (t > 0 || j > 0) ? 1u : 0u);
            if (A_SPLIT) umma_bf16(acc, A_MN ? make_desc_mn_sw128(sa_lo + j * 2048) : make_desc_sw128(sa_lo + j * 32), bh, IDESC, 1u);
          }
#endif
          umma_commit(empty0 + 8 * stage);            // frees the smem stage when these MMAs retire
          if (t == n_kt - 1) umma_commit(tfull0 + 8 * buf);   // accumulator complete
        }
        __syncwarp();
        PROF(2);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (n_kt == 0 && lane == 0) mbar_arrive(tfull0 + 8 * buf);
    }
    PROF_FLUSH(1);
  } else {
    // ================= epilogue (the 4 warps after the MMA warp; warp & 3 = TMEM lane quarter) =================
    const int quarter = warp & 3;
    typename EPI::State est;
    epi.init(est);
    int it = 0;
    PROF_DECL
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
      const int z = tile / mn_tiles, rem = tile - z * mn_tiles;
      const int nt_i = rem / m_tiles, mt_i = rem - nt_i * m_tiles;
      const int m0 = mt_i * 128, n0 = nt_i * NT;
      const int kbeg = CLASSES ? 0 : z * k_chunk, kend = CLASSES ? k_ext : min(k_ext, kbeg + k_chunk);
      const int buf = it & 1;
      if (CLASSES) epi.set_class(z);
      PROF(3);
      mbar_wait(tfull0 + 8 * buf, (it >> 1) & 1);
      PROF(0);
      if (warp == TC_MMA_WARP + 1) TC_TRACE(30);     // accumulator ready
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * Cfg::ACC_COLS;
#pragma unroll
      for (int c = 0; c < NT; c += 16) {
        float v[16];
        tmem_ld16(lane_addr + c, v);
        if (B_SPLIT) {
          float w[16];
          tmem_ld16(lane_addr + NT + c, w);
#pragma unroll
          for (int j = 0; j < 16; j++) v[j] += w[j];
        }
        PROF(1);
        if (warp == TC_MMA_WARP + 1) TC_TRACE(32);
#ifndef XTB_EXP_NO_EPI
        if (kend > kbeg) epi.store16(est, c, row, n0 + c, v);
#endif
        PROF(2);
        if (warp == TC_MMA_WARP + 1) TC_TRACE(33);
      }
      tc_fence_before();
      mbar_arrive(tempty0 + 8 * buf);
      if (warp == TC_MMA_WARP + 1) TC_TRACE(31);     // tile stored
    }
    PROF(3);
    epi.finish(est, epi_sh, NT);
    PROF(4);
    if (warp == TC_MMA_WARP + 1) PROF_FLUSH(2);
  }
  tc_fence_before();
  PROF_STAMP(27);     // this thread's role loop finished
  __syncthreads();
  PROF_STAMP(28);     // all roles finished
  if (warp == TC_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    TC_TRACE(3);
  }
}

// ------------------------------------------------------------------------------------------
// plane maintenance kernels
// ------------------------------------------------------------------------------------------
// fp32 -> bf16 hi / lo planes (weights after an optimiser step or set_weights; any fp32 tensor)
__global__ void split_planes_kernel(const float* __restrict__ src, bf16* __restrict__ hi, long long lo_off, long long n) {
  pdl_wait(); pdl_trigger();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(src + i);
    uint32_t h0, l0, h1, l1;
    split2(v.x, v.y, h0, l0); split2(v.z, v.w, h1, l1);
    *reinterpret_cast<uint2*>(hi + i) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(hi + lo_off + i) = make_uint2(l0, l1);
  } else {
    for (long long j = i; j < n; j++) store_plane1(hi, lo_off, j, src[j]);
  }
}
// uint8 frame decode with the minibatch gather folded in: dst[b, :] = bf16(obs[idx[b], :])   (exact)
__global__ void decode_u8_kernel(const uint8_t* __restrict__ obs, const int32_t* __restrict__ idx, int per_sample,
                                 bf16* __restrict__ dst) {
  pdl_wait(); pdl_trigger();
  int b = blockIdx.y;
  long long srow = idx ? idx[b] : b;
  const uint8_t* s = obs + srow * per_sample;
  bf16* d = dst + (long long)b * per_sample;
  int i = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i + 15 < per_sample) {
    uint4 v = *reinterpret_cast<const uint4*>(s + i);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[8];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      o[2 * q] = pack_bf16((float)(w[q] & 0xff), (float)((w[q] >> 8) & 0xff));
      o[2 * q + 1] = pack_bf16((float)((w[q] >> 16) & 0xff), (float)(w[q] >> 24));
    }
    uint4* dp = reinterpret_cast<uint4*>(d + i);
    dp[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dp[1] = make_uint4(o[4], o[5], o[6], o[7]);
  } else {
    for (int j = i; j < per_sample; j++) d[j] = __float2bfloat16_rn((float)s[j]);
  }
}

// Space-to-depth variant for a first conv layer with stride S = 4 over C = 4 channels (KH, KW, H, W multiples of 4):
// dst[b, Y, X, (dy, dx, c)] = obs[idx[b], 4Y+dy, 4X+dx, c].  The 8x8/4 conv over [H, W, 4] becomes a 2x2/1 conv over
// [H/4, W/4, 64]: every im2col row piece is then one aligned 128-byte line instead of overlapping 64-byte pieces at
// 32-byte offsets (which cost ~17 L2 requests per 512-byte copy instruction).  One thread moves one (y, 4-pixel) unit.
__global__ void decode_u8_s2d_kernel(const uint8_t* __restrict__ obs, const int32_t* __restrict__ idx, int H, int W,
                                     bf16* __restrict__ dst) {
  pdl_wait(); pdl_trigger();
  const int b = blockIdx.y;
  const int wb = W >> 2;                       // 16-byte units (4 pixels x 4 channels) per image row
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= H * wb) return;
  const int y = u / wb, X = u - y * wb;
  const long long per = (long long)H * W * 4;
  const long long srow = idx ? idx[b] : b;
  const uint4 v = *reinterpret_cast<const uint4*>(obs + srow * per + ((long long)y * W + 4 * X) * 4);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[8];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    o[2 * q] = pack_bf16((float)(w[q] & 0xff), (float)((w[q] >> 8) & 0xff));
    o[2 * q + 1] = pack_bf16((float)((w[q] >> 16) & 0xff), (float)(w[q] >> 24));
  }
  bf16* d = dst + (long long)b * per + ((long long)((y >> 2) * wb + X) * 64 + (y & 3) * 16);
  uint4* dp = reinterpret_cast<uint4*>(d);
  dp[0] = make_uint4(o[0], o[1], o[2], o[3]);
  dp[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

}  // namespace tc
}  // namespace xtb
