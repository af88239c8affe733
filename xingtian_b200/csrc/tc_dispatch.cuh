// tc_dispatch.cuh -- host-side selection of tcgen05 kernel instantiations per layer operation.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
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
  static int resident = 0;      // persistent CTAs the whole GPU holds at once (registers can cap what shared memory allows)
  if (!resident) {
    cudaError_t e0 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e0 != cudaSuccess) return e0;
    // ask for the full shared-memory carve-out: the occupancy query otherwise assumes the default split and under-counts
    e0 = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e0 != cudaSuccess) return e0;
    // Resident CTAs per SM from the kernel's real footprint (cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1
    // for these kernels on this driver although two are co-resident): registers are allocated per warp in units
    // of 8 per thread, shared memory per block = dynamic + static + 1 KiB reserved.
    int dev = 0, sms = 0, smem_sm = 0, regs_sm = 0;
    cudaFuncAttributes fa;
    if ((e0 = cudaGetDevice(&dev)) != cudaSuccess) return e0;
    if ((e0 = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e0;
    if ((e0 = cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev)) != cudaSuccess) return e0;
    if ((e0 = cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev)) != cudaSuccess) return e0;
    if ((e0 = cudaFuncGetAttributes(&fa, kern)) != cudaSuccess) return e0;
    const int regs_cta = ((fa.numRegs + 7) / 8 * 8) * 32 * (TC_THREADS / 32);
    const int smem_cta = Cfg::SMEM_BYTES + (int)fa.sharedSizeBytes + 1024;
    int occ = std::min(regs_sm / regs_cta, smem_sm / smem_cta);
    if (getenv("XTB_TC_DEBUG"))
      fprintf(stderr, "[xtb tc] NT=%d A_MN=%d B_MN=%d split=%d%d: %d regs, %d B shared per CTA -> %d CTA/SM (cap %d)\n", NT, (int)A_MN,
              (int)B_MN, (int)A_SPLIT, (int)B_SPLIT, fa.numRegs, smem_cta, occ, Cfg::CTAS_PER_SM);
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    resident = sms * std::min(occ, Cfg::CTAS_PER_SM);
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
  int grid = (int)std::min<long long>(total, resident);
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
