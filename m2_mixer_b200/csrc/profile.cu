// Launch accounting for bench.py: every kernel launch of this library goes through a LaunchScope.
//  - always: a relaxed atomic launch counter (bench.py reports it as `gpu_launches`)
//  - when enabled (m2b200_profile_enable(1)): a cudaEvent pair on the LAUNCHING stream around the launch, so that
//    per-kernel device time can be read back after a synchronize (roofline.achieved is computed from these).
// Profiling is off by default and adds nothing to the hot path but one atomic increment.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/m2b200.h"
#include "common.cuh"
#include "kernels.h"

namespace m2 {
const uint32_t* g_drop_epoch_ptr = nullptr;   // device counter folded into every dropout key (common.cuh drop_key)
namespace {
__global__ void epoch_advance_kernel(uint32_t* e) { *e += 1u; }
std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_enabled{0};
struct Rec { const char* name; cudaEvent_t a, b; };
std::mutex g_mu;
std::vector<Rec> g_recs;
}  // namespace

LaunchScope::LaunchScope(const char* name, cudaStream_t s, int nkernels) : name_(name), stream_(s), start_(nullptr) {
  g_launches.fetch_add(static_cast<unsigned long long>(nkernels), std::memory_order_relaxed);
  if (g_enabled.load(std::memory_order_relaxed)) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) == cudaSuccess) {
      cudaEventRecord(e, s);
      start_ = e;
    }
  }
}

LaunchScope::~LaunchScope() {
  if (!start_) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, stream_);
  std::lock_guard<std::mutex> lk(g_mu);
  g_recs.push_back({name_, static_cast<cudaEvent_t>(start_), e});
}
}  // namespace m2

extern "C" {

void m2b200_set_dropout_epoch_ptr(const void* dev_u32) { m2::g_drop_epoch_ptr = static_cast<const uint32_t*>(dev_u32); }

int m2b200_dropout_epoch_advance(void* dev_u32, void* stream) {
  if (!dev_u32) return M2_ERR_ARG;
  m2::LaunchScope scope("epoch_advance", static_cast<cudaStream_t>(stream));
  m2::epoch_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint32_t*>(dev_u32));
  M2_LAUNCH_CHECK();
  return M2_OK;
}

unsigned long long m2b200_launch_count(void) { return m2::g_launches.load(); }

void m2b200_profile_enable(int on) { m2::g_enabled.store(on ? 1 : 0); }

// Synchronises the recorded events, writes "name,launches,total_ms\n" lines into buf (NUL terminated), clears the log.
// Returns the number of bytes that the full report needs (excluding the NUL).
size_t m2b200_profile_collect(char* buf, size_t cap) {
  std::vector<m2::Rec> recs;
  {
    std::lock_guard<std::mutex> lk(m2::g_mu);
    recs.swap(m2::g_recs);
  }
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto& r : recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += ms;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s,%lld,%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && cap) {
    const size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return out.size();
}

}  // extern "C"
