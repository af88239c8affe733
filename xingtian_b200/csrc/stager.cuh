// stager.cuh -- pageable host memory -> HBM through a pinned ring filled by a small thread pool.
//
// The reference feeds every sess.run() from pageable numpy arrays (feed_dict, xt/model/ppo/ppo.py:104-132,
// xt/model/tf_compat usage in xt/model/model.py:47-65); TF stages them with one host thread.  Here the
// calling thread splits the source into chunks, worker threads memcpy chunks into pinned slots in
// parallel, and the caller issues one cudaMemcpyAsync per run of staged chunks (contiguous in the ring), so the host
// memcpy (the slow leg: ~10 GB/s per core) runs on several cores and overlaps the DMA.
//
// Semantics of stage_h2d(): on return the source buffer has been fully consumed (it may be reused or
// freed) and every copy is enqueued on `stream`, i.e. the same contract as cudaMemcpy from pageable
// memory, minus the stream synchronisation.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace xtb {

class Stager {
 public:
  // Slot size: XTB_STAGE_CHUNK_KB (default 256).  Measured on the C2 inference upload (0.9 MB, profiles/r2_results.md):
  // 256 KiB slots / 4 threads 76 us, 64 KiB slots 96 us (every DMA issue costs the calling thread several us), pinned
  // memory floor 40 us.
  static constexpr size_t RING_BYTES = 12u << 20;
  static constexpr int MAX_SLOTS = 192;

  static Stager& instance() {
    static Stager* s = new Stager();            // leaked on purpose: workers may outlive static destructors
    return *s;
  }

  // returns cudaSuccess or the first CUDA error
  cudaError_t stage_h2d(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return cudaSuccess;
    std::lock_guard<std::mutex> call_lock(call_mu_);       // one staged copy at a time
    cudaError_t err = init();
    if (err != cudaSuccess) return err;
    const size_t n = (bytes + chunk_ - 1) / chunk_;
    release_if_idle();
    const char* s = static_cast<const char*>(src);
    char* d = static_cast<char*>(dst);
    if (n == 1 || workers_.empty()) {                        // small copy: stage on the calling thread
      for (size_t c = 0; c < n; c++) {
        int slot = (int)((seq_ + c) % slots_);
        if ((err = wait_slot(slot)) != cudaSuccess) return err;
        size_t off = c * chunk_, len = bytes - off < chunk_ ? bytes - off : chunk_;
        memcpy(ring_ + (size_t)slot * chunk_, s + off, len);
        if ((err = issue(slot, 1, d + off, len, stream)) != cudaSuccess) return err;
      }
      seq_ += n;
      return cudaSuccess;
    }
    const bool tr = trace_on();
    const auto t_in = tr ? std::chrono::steady_clock::now() : std::chrono::steady_clock::time_point();
    // publish the job, then open it by resetting `next_`
    if (done_.size() < n) done_ = std::vector<std::atomic<int>>(n);
    for (size_t c = 0; c < n; c++) done_[c].store(0, std::memory_order_relaxed);
    const uint64_t g = gen_.load(std::memory_order_relaxed) + 1;
    // close the claim counter first (generation g, index beyond any chunk count): a worker still holding the previous
    // generation can no longer win a CAS against the old counter value while the job fields below change
    next_.store((g << 32) | 0xffffffffu, std::memory_order_release);
    job_src_.store(s, std::memory_order_relaxed); job_bytes_.store(bytes, std::memory_order_relaxed);
    job_chunks_.store(n, std::memory_order_relaxed); job_seq_.store(seq_, std::memory_order_relaxed);
    size_t allowed = 0;
    while (allowed < n && allowed < (size_t)slots_) {         // slots last used by earlier calls
      if ((err = wait_slot((int)((seq_ + allowed) % slots_))) != cudaSuccess) return err;
      allowed++;
    }
    allowed_.store(allowed, std::memory_order_release);
    next_.store(g << 32, std::memory_order_release);      // claim counter = (generation << 32) | chunk index
    gen_.store(g, std::memory_order_release);
    if (sleepers_.load(std::memory_order_acquire) > 0) { std::lock_guard<std::mutex> lk(mu_); cv_.notify_all(); }

    cudaError_t first = cudaSuccess;
    const size_t nw = workers_.size();
    const auto t_pub = tr ? std::chrono::steady_clock::now() : t_in;
    auto t_first = t_pub;
    bool seen_first = false;
    for (size_t c = 0; c < n;) {
      // wait for chunk c; the caller stages chunks itself only while more are unclaimed than the workers can take
      // (or the workers are still asleep): a chunk in hand would delay the copies that are ready to be issued
      while (!done_[c].load(std::memory_order_acquire)) {
        const size_t claimed = (size_t)(next_.load(std::memory_order_acquire) & 0xffffffffu);
        const bool help = sleepers_.load(std::memory_order_acquire) > 0 || n - std::min(claimed, n) > nw;
        if (!(help && work_one(g))) cpu_relax();
        if (allowed < n && allowed < c + slots_) {            // hand out slots whose previous DMA (issued: chunk < c) finished
          int slot = (int)((seq_ + allowed) % slots_);
          if (query_slot(slot) == cudaSuccess) allowed_.store(++allowed, std::memory_order_release);
        }
      }
      if (tr && !seen_first) { t_first = std::chrono::steady_clock::now(); seen_first = true; }
      // one DMA for the run of finished chunks that is contiguous in the ring: c .. e
      size_t e = c;
      while (e + 1 < n && (int)((seq_ + e + 1) % slots_) != 0 && done_[e + 1].load(std::memory_order_acquire)) e++;
      const int slot0 = (int)((seq_ + c) % slots_);
      const size_t off = c * chunk_, end = std::min(bytes, (e + 1) * chunk_);
      const auto t_a = tr ? std::chrono::steady_clock::now() : t_in;
      if (first == cudaSuccess) first = issue(slot0, (int)(e - c + 1), d + off, end - off, stream);
      if (tr) { trace_.api += us(t_a, std::chrono::steady_clock::now()); trace_.dmas++; }
      c = e + 1;
      if (allowed < n && allowed <= c) {                     // nothing else can be staged: block on the oldest DMA
        int s2 = (int)((seq_ + allowed) % slots_);
        cudaError_t e2 = wait_slot(s2);
        if (first == cudaSuccess) first = e2;
        allowed_.store(++allowed, std::memory_order_release);
      }
    }
    seq_ += n;
    if (tr) {
      const auto t_out = std::chrono::steady_clock::now();
      trace_.publish += us(t_in, t_pub); trace_.first_done += us(t_pub, t_first); trace_.issue_loop += us(t_first, t_out);
      if (++trace_.calls == 128) {
        fprintf(stderr, "[stager] %zu chunks/call: publish %.1f us, first chunk staged after %.1f us, issue loop %.1f us (of which "
                "copy/record API %.1f us), %.2f DMAs per call\n", n, trace_.publish / 128, trace_.first_done / 128,
                trace_.issue_loop / 128, trace_.api / 128, trace_.dmas / 128.0);
        trace_ = Trace();
      }
    }
    return first;
  }

  int threads() const { return (int)workers_.size(); }

  // XTB_STAGE_TRACE=1: mean per-call phase times on stderr every 128 multi-chunk calls (diagnosis only)
  struct Trace { double publish = 0, first_done = 0, issue_loop = 0, api = 0; long dmas = 0, calls = 0; };

 private:
  Stager() {}

  static bool trace_on() { static const bool on = getenv("XTB_STAGE_TRACE") != nullptr; return on; }
  static double us(std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::micro>(b - a).count();
  }
  Trace trace_;

  static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }

  cudaError_t init() {
    if (inited_) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    if (n_events_ == 0) {
      if (const char* env = getenv("XTB_STAGE_CHUNK_KB")) {
        long kb = atol(env);
        if (kb >= 64 && kb <= 1024 && (kb & (kb - 1)) == 0) chunk_ = (size_t)kb << 10;
      }
      slots_ = (int)(RING_BYTES / chunk_);
    }
    if (!ring_) {
      e = cudaHostAlloc((void**)&ring_, RING_BYTES, cudaHostAllocDefault);
      if (e != cudaSuccess) { ring_ = nullptr; return e; }
    }
    for (int i = n_events_; i < slots_; i++) {
      e = cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;              // a later call resumes here: nothing half-initialised is used
      used_[i] = false; owner_[i] = i;
      n_events_ = i + 1;
    }
    inited_ = true;
    int nt = 4;
    if (const char* env = getenv("XTB_STAGE_THREADS")) nt = atoi(env);
    unsigned hw = std::thread::hardware_concurrency();
    if (hw && nt > (int)hw - 1) nt = (int)hw - 1;
    if (nt < 0) nt = 0;
    for (int i = 0; i < nt; i++) workers_.emplace_back([this] { worker(); });
    for (auto& t : workers_) t.detach();
    return cudaSuccess;
  }

  // A few queries instead of one wait per slot: when the most recent copy of every stream used since the ring was last
  // idle has completed, every earlier DMA has completed too and the whole ring is free (the common case: the caller
  // synchronised its stream after the previous call).
  void release_if_idle() {
    if (n_ls_ == 0 || ls_overflow_) return;
    for (int k = 0; k < n_ls_; k++)
      if (cudaEventQuery(ev_[ls_ev_[k]]) != cudaSuccess) return;        // cudaErrorNotReady is not sticky
    for (int i = 0; i < slots_; i++) used_[i] = false;
    n_ls_ = 0;
  }

  // A run of slots issued as one DMA shares the event of its last slot (owner_[slot]); that event may be re-recorded
  // by a later run ending in the same slot, which only makes a wait on it more conservative.
  cudaError_t wait_slot(int slot) {
    if (!used_[slot]) return cudaSuccess;
    return cudaEventSynchronize(ev_[owner_[slot]]);
  }
  cudaError_t query_slot(int slot) {
    if (!used_[slot]) return cudaSuccess;
    return cudaEventQuery(ev_[owner_[slot]]);
  }

  // one async copy out of `count` consecutive ring slots starting at `slot`
  cudaError_t issue(int slot, int count, void* dst, size_t len, cudaStream_t stream) {
    cudaError_t e = cudaMemcpyAsync(dst, ring_ + (size_t)slot * chunk_, len, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    const int last = slot + count - 1;
    for (int i = slot; i <= last; i++) { used_[i] = true; owner_[i] = last; }
    int k = 0;
    while (k < n_ls_ && ls_stream_[k] != stream) k++;
    if (k == n_ls_) { if (n_ls_ < MAX_LS) { ls_stream_[n_ls_++] = stream; } else { ls_overflow_ = true; k = 0; } }
    ls_ev_[k] = last;                                                    // most recent DMA of this stream
    return cudaEventRecord(ev_[last], stream);
  }

  // claim and stage one chunk of the open job; false if none is claimable right now
  // A successful CAS on (g, c) proves job g is still open (chunk c was unclaimed), so its fields are stable.
  bool work_one(uint64_t g) {
    uint64_t cur = next_.load(std::memory_order_acquire);
    if ((cur >> 32) != g) return false;
    size_t c = (size_t)(cur & 0xffffffffu);
    if (c >= job_chunks_.load(std::memory_order_relaxed) || c >= allowed_.load(std::memory_order_acquire)) return false;
    if (!next_.compare_exchange_strong(cur, cur + 1, std::memory_order_acq_rel)) return true;   // lost a race: retry
    int slot = (int)((job_seq_.load(std::memory_order_relaxed) + c) % slots_);
    size_t bytes = job_bytes_.load(std::memory_order_relaxed);
    size_t off = c * chunk_, len = bytes - off < chunk_ ? bytes - off : chunk_;
    memcpy(ring_ + (size_t)slot * chunk_, job_src_.load(std::memory_order_relaxed) + off, len);
    done_[c].store(1, std::memory_order_release);
    return true;
  }

  void worker() {
    uint64_t seen = 0;
    for (;;) {
      // spin briefly for the next job (rollout loops call every ~100 us), then sleep
      auto t0 = std::chrono::steady_clock::now();
      int spins = 0;
      while (gen_.load(std::memory_order_acquire) == seen) {
        cpu_relax();
        if ((++spins & 1023) == 0 &&
            std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(300)) {
          std::unique_lock<std::mutex> lk(mu_);
          sleepers_.fetch_add(1, std::memory_order_acq_rel);
          cv_.wait_for(lk, std::chrono::milliseconds(200),
                       [&] { return gen_.load(std::memory_order_acquire) != seen; });
          sleepers_.fetch_sub(1, std::memory_order_acq_rel);
          t0 = std::chrono::steady_clock::now();
        }
      }
      seen = gen_.load(std::memory_order_acquire);
      // drain the job: chunks become claimable as the caller frees slots
      for (;;) {
        if (work_one(seen)) continue;
        uint64_t cur = next_.load(std::memory_order_acquire);
        if ((cur >> 32) != seen || (size_t)(cur & 0xffffffffu) >= job_chunks_.load(std::memory_order_relaxed)) break;
        cpu_relax();
      }
    }
  }

  std::mutex call_mu_, mu_;
  std::condition_variable cv_;
  std::atomic<uint64_t> gen_{0};
  std::atomic<int> sleepers_{0};
  std::atomic<uint64_t> next_{0};
  std::atomic<size_t> allowed_{0};
  std::vector<std::atomic<int>> done_;
  std::atomic<const char*> job_src_{nullptr};
  std::atomic<size_t> job_bytes_{0}, job_chunks_{0}, job_seq_{0};
  size_t seq_ = 0;
  char* ring_ = nullptr;
  bool inited_ = false;
  int n_events_ = 0;
  static constexpr int MAX_LS = 8;
  cudaStream_t ls_stream_[MAX_LS];   // streams with copies issued since the ring was last known idle ...
  int ls_ev_[MAX_LS];                // ... and the event of the most recent copy on each
  int n_ls_ = 0;
  bool ls_overflow_ = false;
  size_t chunk_ = 256u << 10;   // bytes per pinned slot
  int slots_ = 48;              // RING_BYTES / chunk_
  cudaEvent_t ev_[MAX_SLOTS];
  bool used_[MAX_SLOTS];
  int owner_[MAX_SLOTS];
  std::vector<std::thread> workers_;
};

}  // namespace xtb
