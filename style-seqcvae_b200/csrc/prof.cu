#include "prof.cuh"
#include "../../include/sscvae.h"
#include "common.cuh"
#include <map>
#include <string>
#include <vector>

namespace sscvae {
bool g_prof_enabled = false;
struct ProfRec { const char* name; double flops, bytes; cudaEvent_t a, b; };
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
void prof_begin(cudaStream_t s, const char* name, double flops, double bytes) {
  ProfRec r; r.name = name; r.flops = flops; r.bytes = bytes; r.a = get_event(); r.b = get_event();
  cudaEventRecord(r.a, s);
  g_recs.push_back(r);
}
void prof_end(cudaStream_t s) { cudaEventRecord(g_recs.back().b, s); }
}  // namespace sscvae

using namespace sscvae;
extern "C" {
int sscvae_profile_enable(int on) {
  for (ProfRec& r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
  g_recs.clear();
  g_prof_enabled = on != 0;
  return 0;
}
// Writes a JSON object {"class": {"count": n, "ms": t, "flops": f, "bytes": b}, ...} for the launches
// recorded since sscvae_profile_enable(1). Synchronises the device.
int sscvae_profile_report(char* buf, size_t n) {
  REQUIRE(buf && n > 2, "bad buffer");
  CUDA_TRY(cudaDeviceSynchronize());
  struct Agg { long count = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (ProfRec& r : g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
    Agg& a = agg[r.name];
    a.count++; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
  }
  std::string out = "{";
  bool first = true;
  for (auto& kv : agg) {
    char tmp[256];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"count\": %ld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}", first ? "" : ", ",
             kv.first.c_str(), kv.second.count, kv.second.ms, kv.second.flops, kv.second.bytes);
    out += tmp;
    first = false;
  }
  out += "}";
  if (out.size() + 1 > n) { set_error("profile report buffer too small"); return SSCVAE_ERR_WORKSPACE; }
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}
}
