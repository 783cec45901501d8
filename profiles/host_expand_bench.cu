// Host-side microbenchmark of the observation wire expansion of mbe_step_host (no GPU involved):
//   nvcc -O3 -std=c++17 -o /tmp/hxb profiles/host_expand_bench.cu -lpthread && /tmp/hxb <threads> <ma 0|1> [envs]
// Prints ms per step for expanding E envs of the medium shape (15 UEs x 4 BSs) in 8 windows, beside a
// single-thread memcpy of the same output size (the box's per-core copy bandwidth).
#include "../mobile_env_gan_b200/csrc/mbe_host_wire.cuh"
#include <chrono>
#include <cstdio>
#include <cstdlib>
using namespace mbe;
int main(int argc, char** argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 8, ma = argc > 2 ? atoi(argv[2]) : 0;
  const int E = argc > 3 ? atoi(argv[3]) : (ma ? 131072 : 65536), windows = 8;
  WireShape w;
  w.U = 15, w.B = 4, w.F = ma ? 17 : 9, w.MW = 1, w.W = ma ? 2 : 1, w.ma = ma;
  const size_t wb = (size_t)E * w.bytes_per_env(), ob = (size_t)E * w.U * w.F * 4;
  unsigned char* wire = (unsigned char*)aligned_alloc(4096, (wb + 4095) / 4096 * 4096);
  float* obs = (float*)aligned_alloc(4096, ob);
  float* obs2 = (float*)aligned_alloc(4096, ob);
  for (size_t i = 0; i < wb / 4; ++i) ((float*)wire)[i] = (float)(rand() % 4) / 3.f;
  memset(obs, 0, ob), memset(obs2, 1, ob);
  ExpandCrew crew(threads - 1);
  const int per = E / windows;
  for (int rep = 0; rep < 3; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    const int iters = 20;
    for (int it = 0; it < iters; ++it) {
      const int chunks = 4 * threads;
      crew.begin(windows, chunks, [=](int wi, int c) {
        const unsigned char* src = wire + (size_t)wi * per * w.bytes_per_env();
        float* dst = obs + (size_t)wi * per * w.U * w.F;
        const int lo = (int)((long long)per * c / chunks), hi = (int)((long long)per * (c + 1) / chunks);
        wire_expand_any(src, dst, per, lo, hi, w);
      });
      for (int wi = 0; wi < windows; ++wi) crew.publish(wi + 1);
      crew.finish();
    }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / iters;
    printf("threads %d ma %d: expand %.3f ms/step (%.1f GB/s out, wire %.1f MB -> obs %.1f MB)\n", threads, ma, dt * 1e3,
           ob / dt / 1e9, wb / 1e6, ob / 1e6);
    t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < iters; ++it) memcpy(obs2, obs, ob);
    dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / iters;
    printf("  1-thread memcpy of the output: %.3f ms (%.1f GB/s)\n", dt * 1e3, ob / dt / 1e9);
  }
  return 0;
}
