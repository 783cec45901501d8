// Compact wire format of the observation for the host-buffer entry point (mbe_step_host).
//
// An FP32 observation row is mostly redundant: B of its F columns are a one-hot connection mask,
// and the two multi-agent broadcast blocks repeat per-BS values of the env in every row.  PCIe is the
// bottleneck of the host path (profiles/README.md), so the rows cross it packed and are expanded into
// the caller's FP32 buffer by host threads, pipelined per env window:
//
//   vals  f32 [n, U, B+1]   snr / max snr per BS, own utility            (copied verbatim)
//   bits  u32 [n, U, W]     W = MW (central) or 2 MW (multi-agent): connection mask words, then
//                           the words of "row shows BS b" (see below)
//   bsu   f32 [n, B]        multi-agent only: broadcast BS utilities (allStationUtilities, base.py:438-447)
//
// The packer is a pure function of the observation tensor the step kernels wrote, and the expansion
// reproduces that tensor bit for bit (tests compare it with the device tensor):
//   * one-hot columns come back from the mask bits;
//   * a row of an inactive UE is all zeros; an active row holds snr / max snr == 1.0f for its best BS,
//     so "no ratio equals 1.0f" identifies the inactive rows;
//   * multi-agent column 2B+1+b is bsu[b] when BS b is connectable from the UE's position and -1
//     otherwise, column 3B+1+b is cnt[b] / max(1, sum of the connectable cnt) resp. 0.  "Shows BS b" =
//     (column 2B+1+b != -1 or column 3B+1+b != 0); where it is unset both columns are (-1, 0) whatever the
//     connectability was.  cnt[b] is the number of one-hot bits of the env, the quotient is IEEE FP32 on
//     both sides (1.0f / x, then one multiply).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace mbe {

struct WireShape {
  int U, B, F, MW, W, ma;
  __host__ __device__ size_t vals_per_env() const { return (size_t)U * (B + 1); }        // floats
  __host__ __device__ size_t bits_per_env() const { return (size_t)U * W; }              // words
  __host__ __device__ size_t bsu_per_env() const { return ma ? (size_t)B : 0; }          // floats
  __host__ __device__ size_t bytes_per_env() const { return 4 * (vals_per_env() + bits_per_env() + bsu_per_env()); }
};

// One thread per (env, UE) row of the window [first, first + n).  `wire` points at the window's block:
// vals [n,U,B+1] | bits [n,U,W] | bsu [n,B].
__global__ void __launch_bounds__(256) wire_pack_kernel(const float* __restrict__ obs, unsigned char* __restrict__ wire,
                                                        int first, int n, WireShape w) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n * w.U) return;
  const int e = row / w.U;
  const float* src = obs + ((size_t)first * w.U + row) * w.F;
  float* vals = reinterpret_cast<float*>(wire) + (size_t)row * (w.B + 1);
  uint32_t* bits = reinterpret_cast<uint32_t*>(wire + 4 * (size_t)n * w.vals_per_env()) + (size_t)row * w.W;
  float* bsu = reinterpret_cast<float*>(wire + 4 * (size_t)n * (w.vals_per_env() + w.bits_per_env())) + (size_t)e * w.B;
  const int B = w.B;
  for (int wd = 0; wd < w.MW; ++wd) {
    uint32_t m = 0, sh = 0;
    for (int b = 32 * wd; b < min(B, 32 * wd + 32); ++b) {
      if (src[b] != 0.0f) m |= 1u << (b & 31);
      if (w.ma) {
        const float u = src[2 * B + 1 + b];
        if (u != -1.0f || src[3 * B + 1 + b] != 0.0f) {
          sh |= 1u << (b & 31);
          bsu[b] = u;  // the same value from every row that shows b
        }
      }
    }
    bits[wd] = m;
    if (w.ma) bits[w.MW + wd] = sh;
  }
  for (int b = 0; b <= B; ++b) vals[b] = src[B + b];
}

// Expands the wire block of a window into FP32 rows.  Envs [lo, hi) of the window (window-local ids).
template <int CB>  // compile-time B for the scenario shapes, 0 = runtime
inline void wire_expand(const unsigned char* wire, float* obs, int n, int lo, int hi, const WireShape& w) {
  const int U = w.U, B = CB ? CB : w.B, F = w.F, MW = w.MW, W = w.W;
  const float* vals_all = reinterpret_cast<const float*>(wire);
  const uint32_t* bits_all = reinterpret_cast<const uint32_t*>(wire + 4 * (size_t)n * w.vals_per_env());
  const float* bsu_all = reinterpret_cast<const float*>(wire + 4 * (size_t)n * (w.vals_per_env() + w.bits_per_env()));
  float cntf[64];
  for (int e = lo; e < hi; ++e) {
    const uint32_t* bits = bits_all + (size_t)e * U * W;
    const float* bsu = bsu_all + (size_t)e * B;
    if (w.ma) {  // |connections(b)| of the env from the one-hot bits
      int cnt[64] = {0};
      for (int u = 0; u < U; ++u)
        for (int b = 0; b < B; ++b) cnt[b] += (bits[(size_t)u * W + (b >> 5)] >> (b & 31)) & 1u;
      for (int b = 0; b < B; ++b) cntf[b] = (float)cnt[b];
    }
    for (int u = 0; u < U; ++u) {
      const float* vals = vals_all + ((size_t)e * U + u) * (B + 1);
      const uint32_t* m = bits + (size_t)u * W;
      float* row = obs + ((size_t)e * U + u) * F;
      bool active = false;
      for (int b = 0; b < B; ++b) {
        row[b] = ((m[b >> 5] >> (b & 31)) & 1u) ? 1.0f : 0.0f;
        row[B + b] = vals[b];
        active |= vals[b] == 1.0f;
      }
      row[2 * B] = vals[B];
      if (w.ma) {
        if (!active) {
          for (int b = 0; b < 2 * B; ++b) row[2 * B + 1 + b] = 0.0f;
          continue;
        }
        float tsum = 0.0f;
        for (int b = 0; b < B; ++b)
          if ((m[MW + (b >> 5)] >> (b & 31)) & 1u) tsum += cntf[b];
        const float inv = 1.0f / std::max(1.0f, tsum);
        for (int b = 0; b < B; ++b) {
          const bool sh = (m[MW + (b >> 5)] >> (b & 31)) & 1u;
          row[2 * B + 1 + b] = sh ? bsu[b] : -1.0f;
          row[3 * B + 1 + b] = sh ? cntf[b] * inv : 0.0f;
        }
      }
    }
  }
}

inline void wire_expand_any(const unsigned char* wire, float* obs, int n, int lo, int hi, const WireShape& w) {
  switch (w.B) {
    case 3: return wire_expand<3>(wire, obs, n, lo, hi, w);
    case 4: return wire_expand<4>(wire, obs, n, lo, hi, w);
    case 13: return wire_expand<13>(wire, obs, n, lo, hi, w);
    default: return wire_expand<0>(wire, obs, n, lo, hi, w);
  }
}

// A small persistent pool: run(n_chunks, fn) hands chunk ids to the workers (and the caller);
// wait() blocks until the batch is done.  One batch in flight at a time per pool user is enough here:
// batches are queued in order.
class WorkerPool {
 public:
  explicit WorkerPool(int threads) {
    for (int i = 0; i < threads; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // enqueue `chunks` calls fn(0..chunks-1); returns immediately
  void submit(int chunks, std::function<void(int)> fn) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      batches_.push_back(Batch{std::move(fn), chunks, 0, 0});
      pending_ += chunks;
    }
    cv_.notify_all();
  }
  // the caller helps, then waits until every submitted chunk has run
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    while (pending_ > 0) {
      if (!take_and_run(lk)) done_cv_.wait(lk, [this] { return pending_ == 0 || has_work(); });
    }
    batches_.clear();
  }

 private:
  struct Batch {
    std::function<void(int)> fn;
    int chunks, next, finished;
  };
  bool has_work() const {
    for (const Batch& b : batches_)
      if (b.next < b.chunks) return true;
    return false;
  }
  // runs one chunk if any is unclaimed; lk is held on entry and exit
  bool take_and_run(std::unique_lock<std::mutex>& lk) {
    for (size_t i = 0; i < batches_.size(); ++i) {
      Batch& b = batches_[i];
      if (b.next < b.chunks) {
        const int id = b.next++;
        std::function<void(int)>& fn = b.fn;
        lk.unlock();
        fn(id);
        lk.lock();
        batches_[i].finished++;
        if (--pending_ == 0) done_cv_.notify_all();
        return true;
      }
    }
    return false;
  }
  void loop() {
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_.wait(lk, [this] { return stop_ || has_work(); });
      if (stop_) return;
      while (take_and_run(lk)) {
      }
    }
  }
  std::vector<std::thread> workers_;
  std::deque<Batch> batches_;  // deque: submit() must not move a batch a worker is running
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  int pending_ = 0;
  bool stop_ = false;
};

}  // namespace mbe
