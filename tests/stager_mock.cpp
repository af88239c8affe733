// Host-only harness for xingtian_b200/csrc/stager.cuh: a fake CUDA runtime whose copies and events complete on an
// asynchronous "DMA thread", so slot reuse, chunk claiming and source-reuse semantics are exercised without a GPU.
#include <cuda_runtime.h>
#include <deque>
#include <mutex>
#include <thread>
#include <atomic>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <unistd.h>
struct Ev { std::atomic<uint64_t> target{0}; };
struct Op { int kind; void* d; const void* s; size_t n; Ev* ev; uint64_t id; };
static std::atomic<long> n_copies{0};
static std::deque<Op> q; static std::mutex qm; static std::atomic<uint64_t> done_id{0}, next_id{0};
static void dma() { for (;;) { Op o; { std::lock_guard<std::mutex> l(qm); if (q.empty()) { o.kind = -1; } else { o = q.front(); q.pop_front(); } }
  if (o.kind < 0) { usleep(5); continue; }
  if (o.kind == 0) { usleep((o.id % 37 == 0) ? 200 : 3); memcpy(o.d, o.s, o.n); n_copies++; }
  done_id.store(o.id); } }
static void ensure() { static bool st = false; if (!st) { st = true; std::thread(dma).detach(); } }
extern "C" {
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = malloc(n); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t) new Ev(); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { ensure(); std::lock_guard<std::mutex> l(qm); q.push_back({0, d, s, n, nullptr, ++next_id}); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { ensure(); std::lock_guard<std::mutex> l(qm); uint64_t id = ++next_id; ((Ev*)e)->target.store(id); q.push_back({1, 0, 0, 0, (Ev*)e, id}); return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t e) { return done_id.load() >= ((Ev*)e)->target.load() ? cudaSuccess : cudaErrorNotReady; }
cudaError_t cudaEventSynchronize(cudaEvent_t e) { while (cudaEventQuery(e) != cudaSuccess) usleep(1); return cudaSuccess; }
}
#include "stager.cuh"
int main() {
  std::mt19937_64 rng(1);
  auto& st = xtb::Stager::instance();
  size_t ring = xtb::Stager::RING_BYTES;
  const size_t chunk = getenv("XTB_STAGE_CHUNK_KB") ? (size_t)atol(getenv("XTB_STAGE_CHUNK_KB")) << 10 : (256u << 10);
  std::vector<size_t> sizes = {1, 100, chunk, chunk + 1, 903168, 3612672, ring - 1, ring + 77, 4 * ring + 12345};
  for (int it = 0; it < 150; it++) {
    size_t n = it < (int)sizes.size() ? sizes[it] : (rng() % (it % 7 == 0 ? 3 * ring : 2000000)) + 1;
    std::vector<unsigned char> src(n), dst(n, 0), keep;
    for (size_t i = 0; i < n; i += 1 + (n > 100000 ? 61 : 0)) src[i] = (unsigned char)rng();
    keep = src;
    if (st.stage_h2d(dst.data(), src.data(), n, 0) != cudaSuccess) { printf("err\n"); return 1; }
    memset(src.data(), 0xEE, n);        // source reusable immediately
    while (done_id.load() < next_id.load()) usleep(10);
    if (memcmp(dst.data(), keep.data(), n)) { printf("MISMATCH it=%d n=%zu\n", it, n); return 1; }
    if (it % 50 == 0) { printf("it %d ok (n=%zu, threads=%d)\n", it, n, st.threads()); usleep(it % 100 == 0 ? 2000 : 0); }
  }
  printf("dma copies issued: %ld\n", n_copies.load());
  printf("all ok\n"); return 0;
}
