// Microbenchmark: how many L2 requests / how much time does it take to stage 128 rows x 128 B tiles from global
// memory (L2 resident) into shared memory with different instructions and destination layouts?
//   mode 0: cp.async 16 B, 8 lanes per row, identity destination      (row r, chunk c) -> r*128 + c*16
//   mode 1: cp.async 16 B, 8 lanes per row, 128B-swizzled destination (c ^ (r & 7))
//   mode 2: ld.global.nc.v4 + st.shared.v4, swizzled destination
//   mode 3: cp.async 16 B, lanes along rows (one row per lane), core-matrix destination c*2048 + r*16
//   mode 4: cp.async.bulk 128 B per row (one lane per row), identity destination
//   mode 5: cp.async.bulk 16 KB per tile (contiguous tile), identity destination
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_modes gather_modes.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int STAGES = 4, TILE = 16384;

template <int MODE>
__global__ void __launch_bounds__(256) k(const uint8_t* __restrict__ src, int tiles, int pitch, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bars[STAGES];
  const int tid = threadIdx.x;
  const uint32_t base = smem_u32(sm);
  if (MODE >= 4) {
    if (tid == 0) { for (int s = 0; s < STAGES; s++) mbar_init(smem_u32(&bars[s]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
  }
  unsigned acc = 0;
  int it = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
    const int stage = it % STAGES;
    const uint8_t* g = src + (size_t)tile * 128 * pitch;
    const uint32_t d = base + stage * TILE;
    if (MODE <= 1) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int r = (tid >> 3) + 32 * i, c = tid & 7;
        cp16(d + r * 128 + ((MODE == 1 ? (c ^ (r & 7)) : c) << 4), g + (size_t)r * pitch + c * 16);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
    } else if (MODE == 2) {
      uint4 v[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int r = (tid >> 3) + 32 * i, c = tid & 7;
        v[i] = __ldg(reinterpret_cast<const uint4*>(g + (size_t)r * pitch + c * 16));
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int r = (tid >> 3) + 32 * i, c = tid & 7;
        *reinterpret_cast<uint4*>(sm + stage * TILE + r * 128 + ((c ^ (r & 7)) << 4)) = v[i];
      }
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int s = tid + 256 * i; int r = s & 127, c = s >> 7;
        cp16(d + c * 2048 + r * 16, g + (size_t)r * pitch + c * 16);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
    } else if (MODE == 4) {
      const uint32_t bar = smem_u32(&bars[stage]);
      if (it >= STAGES) mbar_wait(bar, ((it / STAGES) - 1) & 1);     // previous use of this stage landed
      __syncthreads();
      if (tid == 0) mbar_expect(bar, TILE);
      __syncthreads();
      if (tid < 128) bulk(d + tid * 128, g + (size_t)tid * pitch, 128, bar);
    } else {
      const uint32_t bar = smem_u32(&bars[stage]);
      if (it >= STAGES) mbar_wait(bar, ((it / STAGES) - 1) & 1);
      __syncthreads();
      if (tid == 0) { mbar_expect(bar, TILE); bulk(d, g, TILE, bar); }
    }
  }
  if (MODE <= 1 || MODE == 3) asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (MODE >= 4) {
    for (int s = 0; s < STAGES && s < it; s++) { int last = ((it - 1 - s) / STAGES); (void)last; }
    // wait for every stage's last use
    for (int j = (it > STAGES ? it - STAGES : 0); j < it; j++) mbar_wait(smem_u32(&bars[j % STAGES]), (j / STAGES) & 1);
  }
  __syncthreads();
  acc += sm[tid * 16];
  if (acc == 0xffffffffu) *sink = acc;
}

template <int MODE>
void run(const uint8_t* src, int tiles, int pitch, unsigned* sink, const char* name) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * TILE);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 2; w++) k<MODE><<<296, 256, STAGES * TILE>>>(src, tiles, pitch, sink);
  cudaEventRecord(a);
  for (int w = 0; w < 5; w++) k<MODE><<<296, 256, STAGES * TILE>>>(src, tiles, pitch, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  double gb = (double)tiles * TILE / 1e9;
  printf("mode %d %-44s pitch %4d: %7.1f us  %7.1f GB/s  %s\n", MODE, name, pitch, ms / 5 * 1e3, gb / (ms / 5 * 1e-3), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char** argv) {
  int tiles = 4096;                        // 64 MiB at pitch 128
  uint8_t* src; unsigned* sink;
  cudaMalloc(&src, (size_t)tiles * 128 * 256 + 4096); cudaMalloc(&sink, 4);
  cudaMemset(src, 1, (size_t)tiles * 128 * 256 + 4096);
  for (int pitch : {128, 256}) {
    if (pitch == 256) tiles = 2048;
    run<0>(src, tiles, pitch, sink, "cp.async 8 lanes/row, identity dst");
    run<1>(src, tiles, pitch, sink, "cp.async 8 lanes/row, swizzled dst");
    run<2>(src, tiles, pitch, sink, "ldg.v4 + sts.v4, swizzled dst");
    run<3>(src, tiles, pitch, sink, "cp.async lane per row, core-matrix dst");
    run<4>(src, tiles, pitch, sink, "cp.async.bulk 128 B per row");
    if (pitch == 128) run<5>(src, tiles, pitch, sink, "cp.async.bulk 16 KB per tile");
  }
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
