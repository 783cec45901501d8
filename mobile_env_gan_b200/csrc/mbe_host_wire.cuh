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
#include <cstring>
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
// The scenario shapes (B <= 32 known at compile time) take the branch-free path below: 0/1 floats are
// built from the mask word in integer registers, the ratio columns are a fixed-size copy.
template <int CB>  // compile-time B for the scenario shapes, 0 = runtime
inline void wire_expand(const unsigned char* __restrict__ wire, float* __restrict__ obs, int n, int lo, int hi,
                        const WireShape& w) {
  const int U = w.U, B = CB ? CB : w.B, F = w.F, MW = w.MW, W = w.W;
  const float* __restrict__ vals_all = reinterpret_cast<const float*>(wire);
  const uint32_t* __restrict__ bits_all = reinterpret_cast<const uint32_t*>(wire + 4 * (size_t)n * w.vals_per_env());
  const float* __restrict__ bsu_all =
      reinterpret_cast<const float*>(wire + 4 * (size_t)n * (w.vals_per_env() + w.bits_per_env()));
  constexpr uint32_t kOne = 0x3f800000u;  // 1.0f
  if (CB > 0 && CB <= 32 && !w.ma) {      // central rows are independent of their env: one flat row loop
    const size_t r0 = (size_t)lo * U, r1 = (size_t)hi * U;
    const float* __restrict__ v = vals_all + r0 * (CB + 1);
    uint32_t* __restrict__ row = reinterpret_cast<uint32_t*>(obs + r0 * (2 * CB + 1));
    for (size_t r = r0; r < r1; ++r, v += CB + 1, row += 2 * CB + 1) {
      const uint32_t m = bits_all[r];
      for (int b = 0; b < CB; ++b) row[b] = (0u - ((m >> b) & 1u)) & kOne;
      std::memcpy(row + CB, v, 4 * (CB + 1));
    }
    return;
  }
  if (CB > 0 && CB <= 32 && w.ma) {  // multi-agent, one mask word: per-env counts, then branch-light rows
    constexpr int FF = 4 * CB + 1;
    for (int e = lo; e < hi; ++e) {
      const uint32_t* __restrict__ bits = bits_all + (size_t)e * U * 2;
      const float* __restrict__ bsu = bsu_all + (size_t)e * CB;
      float cnt[CB > 0 ? CB : 1];
      for (int b = 0; b < CB; ++b) {
        int c = 0;
        for (int u = 0; u < U; ++u) c += (bits[2 * u] >> b) & 1u;
        cnt[b] = (float)c;
      }
      const float* __restrict__ v = vals_all + (size_t)e * U * (CB + 1);
      uint32_t* __restrict__ row = reinterpret_cast<uint32_t*>(obs + (size_t)e * U * FF);
      for (int u = 0; u < U; ++u, v += CB + 1, row += FF) {
        const uint32_t m = bits[2 * u], sh = bits[2 * u + 1];
        bool active = false;
        for (int b = 0; b < CB; ++b) {
          row[b] = (0u - ((m >> b) & 1u)) & kOne;
          active |= v[b] == 1.0f;
        }
        std::memcpy(row + CB, v, 4 * (CB + 1));
        float* __restrict__ tail = reinterpret_cast<float*>(row + 2 * CB + 1);
        if (!active) {
          std::memset(tail, 0, 8 * CB);
          continue;
        }
        float tsum = 0.0f;
        for (int b = 0; b < CB; ++b) tsum += ((sh >> b) & 1u) ? cnt[b] : 0.0f;  // adding +0.0f keeps the sum's bits
        const float inv = 1.0f / std::max(1.0f, tsum);
        for (int b = 0; b < CB; ++b) {
          const bool s = (sh >> b) & 1u;
          tail[b] = s ? bsu[b] : -1.0f;
          tail[CB + b] = s ? cnt[b] * inv : 0.0f;
        }
      }
    }
    return;
  }
  float cntf[64];
  for (int e = lo; e < hi; ++e) {
    const uint32_t* __restrict__ bits = bits_all + (size_t)e * U * W;
    const float* __restrict__ bsu = bsu_all + (size_t)e * B;
    if (w.ma) {  // |connections(b)| of the env from the one-hot bits
      int cnt[64] = {0};
      for (int u = 0; u < U; ++u)
        for (int b = 0; b < B; ++b) cnt[b] += (bits[(size_t)u * W + (b >> 5)] >> (b & 31)) & 1u;
      for (int b = 0; b < B; ++b) cntf[b] = (float)cnt[b];
    }
    for (int u = 0; u < U; ++u) {
      const float* __restrict__ vals = vals_all + ((size_t)e * U + u) * (B + 1);
      const uint32_t* __restrict__ m = bits + (size_t)u * W;
      float* __restrict__ row = obs + ((size_t)e * U + u) * F;
      uint32_t* __restrict__ rowi = reinterpret_cast<uint32_t*>(row);
      bool active = false;
      for (int b = 0; b < B; ++b) {
        rowi[b] = (0u - ((m[b >> 5] >> (b & 31)) & 1u)) & kOne;
        active |= vals[b] == 1.0f;
      }
      std::memcpy(row + B, vals, 4 * (size_t)(B + 1));
      if (w.ma) {
        if (!active) {
          std::memset(row + 2 * B + 1, 0, 8 * (size_t)B);
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

// Host threads that expand the windows of one mbe_step_host call ("session").  Between calls the
// workers sleep on a condition variable; begin() wakes them BEFORE the caller enqueues the GPU work, so
// the wake-up latency (tens of microseconds per futex wake on a virtualised host -- paid once per
// window, it cost more than the expansion itself) hides behind the launches.  Inside a session nothing
// sleeps or locks: the caller publishes window w once its bytes have landed (publish), workers spin on
// that counter and claim chunks of the window with one atomic add each.  finish() makes the caller the
// last worker and returns when every chunk has run; abort() releases the workers without work (error
// paths).
class ExpandCrew {
 public:
  static constexpr int kMaxWindows = 64;
  using Fn = std::function<void(int window, int chunk)>;

  explicit ExpandCrew(int workers) {
    for (int i = 0; i < workers; ++i) threads_.emplace_back([this] { loop(); });
  }
  ~ExpandCrew() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int size() const { return (int)threads_.size(); }

  void begin(int windows, int chunks, Fn fn) {
    {
      // workers enter a session only while holding mu_, so once nobody is inside and the lock is held
      // the session state can be rewritten; late wakers of an older session then join this one
      std::unique_lock<std::mutex> lk(mu_);
      while (inside_.load(std::memory_order_acquire) != 0) {
        lk.unlock();
        relax();
        lk.lock();
      }
      fn_ = std::move(fn);
      windows_ = std::min(windows, kMaxWindows), chunks_ = chunks;
      for (int w = 0; w < windows_; ++w) next_[w].store(0, std::memory_order_relaxed);
      published_.store(0, std::memory_order_relaxed);
      done_.store(0, std::memory_order_relaxed);
      abort_.store(false, std::memory_order_release);
      ++session_;
    }
    cv_.notify_all();
    open_ = true;
  }
  void publish(int windows_ready) { published_.store(windows_ready, std::memory_order_release); }
  // the caller becomes the last worker; returns when every chunk of every window has run
  void finish() {
    work();
    while (done_.load(std::memory_order_acquire) < windows_ * chunks_) relax();
    open_ = false;
  }
  // error paths: no chunk starts after this, and it returns only when no worker touches the buffers
  void abort() {
    if (!open_) return;
    abort_.store(true, std::memory_order_release);
    while (inside_.load(std::memory_order_acquire) != 0) relax();
    open_ = false;
  }
  bool open() const { return open_; }

 private:
  static void relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }
  void work() {
    for (int w = 0; w < windows_; ++w) {
      while (published_.load(std::memory_order_acquire) <= w) {
        if (abort_.load(std::memory_order_acquire)) return;
        relax();
      }
      for (;;) {
        if (abort_.load(std::memory_order_acquire)) return;
        const int c = next_[w].fetch_add(1, std::memory_order_relaxed);
        if (c >= chunks_) break;
        fn_(w, c);
        done_.fetch_add(1, std::memory_order_release);
      }
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || session_ != seen; });
        if (stop_) return;
        seen = session_;
        inside_.fetch_add(1, std::memory_order_acq_rel);
      }
      work();
      inside_.fetch_sub(1, std::memory_order_acq_rel);
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_;
  uint64_t session_ = 0;
  bool stop_ = false, open_ = false;
  Fn fn_;
  int windows_ = 0, chunks_ = 0;
  std::atomic<int> next_[kMaxWindows];
  std::atomic<int> published_{0}, done_{0}, inside_{0};
  std::atomic<bool> abort_{false};
};

}  // namespace mbe
