// Opt-in per-launch timing: when enabled every launcher brackets its kernel with a pair of CUDA
// events on the launching stream and attributes the duration (plus the launch's algorithmic FLOPs /
// bytes) to a category.  bench.py uses it to report the dominant kernel's achieved rate measured
// inside real steps.  Disabled (zero overhead beyond one branch) by default; never enabled while a
// CUDA graph is being captured.
#include <vector>

#include "host_util.h"
#include "kernels.h"

namespace b2 {

namespace {
struct Rec { cudaEvent_t a, b; int cat; double flops, bytes; };
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  B2_CUDA(cudaEventCreate(&e));
  return e;
}
}  // namespace

void prof_enable(bool on) { g_on = on; }
bool prof_enabled() { return g_on; }

ProfScope::ProfScope(int cat, double flops, double bytes, cudaStream_t s) : stream(s), idx(-1) {
  if (!g_on) return;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return;
  Rec r{get_event(), get_event(), cat, flops, bytes};
  cudaEventRecord(r.a, s);
  g_recs.push_back(r);
  idx = (int)g_recs.size() - 1;
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(g_recs[idx].b, stream);
}

void prof_collect(double* ms, double* flops, double* bytes, long long* launches) {
  for (int c = 0; c < PC_COUNT; ++c) { ms[c] = 0; flops[c] = 0; bytes[c] = 0; launches[c] = 0; }
  for (auto& r : g_recs) {
    cudaEventSynchronize(r.b);
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms[r.cat] += t; flops[r.cat] += r.flops; bytes[r.cat] += r.bytes; launches[r.cat] += 1;
    }
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
}

}  // namespace b2
