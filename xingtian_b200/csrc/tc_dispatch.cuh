// tc_dispatch.cuh -- host-side selection of tcgen05 kernel instantiations per layer operation.
#pragma once
#include <algorithm>
#include "gemm_tc.cuh"

namespace xtb {
namespace tc {

// classes > 0: the kernel runs `classes` independent GEMMs over the full K range (parity classes of a strided
// transposed conv, see SrcDgradS) instead of a K split
template <int NT, int KT, bool A_MN, bool B_MN, bool A_SPLIT, bool B_SPLIT, class AS, class BS, class EP, bool CLASSES = false>
static cudaError_t launch_tc(const AS& a, const BS& b, const EP& e, int m_ext, int n_ext, int k_ext, int ksplit,
                             cudaStream_t st, int classes = 0) {
  using Cfg = TcCfg<NT, KT, A_MN, B_MN, A_SPLIT, B_SPLIT>;
  auto kern = tc_gemm_kernel<NT, KT, A_MN, B_MN, A_SPLIT, B_SPLIT, CLASSES, AS, BS, EP>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e0 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e0 != cudaSuccess) return e0;
    attr_set = true;
  }
  if (ksplit < 1) ksplit = 1;
  int kc = (k_ext + ksplit - 1) / ksplit;
  kc = (kc + KT - 1) / KT * KT;
  // K-major conv operands cache the column groups of one tile's K range in shared memory
  if ((AS::TABLE && !A_MN) || (BS::TABLE && !B_MN)) {
    while (kc / 8 > TC_MAXCOLS) { ksplit++; kc = ((k_ext + ksplit - 1) / ksplit + KT - 1) / KT * KT; }
  }
  int m_tiles = (m_ext + 127) / 128, n_tiles = (n_ext + NT - 1) / NT, k_tiles = (k_ext + kc - 1) / kc;
  if (CLASSES) { if (kc < k_ext || k_ext / 8 > TC_MAXCOLS) return cudaErrorInvalidValue; k_tiles = classes; }
  long long total = (long long)m_tiles * n_tiles * k_tiles;
  int grid = (int)std::min<long long>(total, 148LL * Cfg::CTAS_PER_SM);
  XLAUNCH(kern, grid, TC_THREADS, Cfg::SMEM_BYTES, st, a, b, e, m_ext, n_ext, k_ext, kc, m_tiles, n_tiles, (int)total);
  return cudaPeekAtLastError();
}

// pick the N tile: smallest of {16,32,64,128,256} covering n (or 256 tiles)
#define XTB_TC_SWITCH_NT(n, CALL)            \
  do {                                       \
    if ((n) <= 16) { CALL(16); }             \
    else if ((n) <= 32) { CALL(32); }        \
    else if ((n) <= 64) { CALL(64); }        \
    else if ((n) <= 128) { CALL(128); }      \
    else { CALL(256); }                      \
  } while (0)

static inline int pick_ksplit(long long ctas, int k_ext, int kt, int target) {
  if (ctas >= target) return 1;
  int ks = (int)((target + ctas - 1) / ctas);
  int mx = (k_ext + kt - 1) / kt;
  if (ks > mx) ks = mx;
  return ks < 1 ? 1 : ks;
}

}  // namespace tc
}  // namespace xtb
