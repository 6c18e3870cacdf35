// Optional per-launch instrumentation (off by default): CUDA events on the launching stream around every
// kernel launch of the library, aggregated per kernel class with the algorithmic FLOPs / bytes the call
// site declares. bench.py uses it for the roofline object and the per-kernel share of a step.
#pragma once
#include <cuda_runtime.h>

namespace sscvae {
extern bool g_prof_enabled;
void prof_begin(cudaStream_t s, const char* name, double flops, double bytes);
void prof_end(cudaStream_t s);
struct ProfScope {
  cudaStream_t s; bool on;
  ProfScope(cudaStream_t st, const char* name, double flops, double bytes) : s(st), on(g_prof_enabled) {
    if (on) prof_begin(s, name, flops, bytes);
  }
  ~ProfScope() { if (on) prof_end(s); }
};
}  // namespace sscvae
#define PROF_SCOPE(stream, name, flops, bytes) sscvae::ProfScope _prof_scope((stream), (name), (double)(flops), (double)(bytes))
