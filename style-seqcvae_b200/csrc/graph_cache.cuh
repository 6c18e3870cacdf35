// CUDA-graph replay of a whole C-ABI call (training forward, BPTT, decode).
//
// A training step is ~580 kernel launches of 3-25 us each on one stream; what is left between them is launch
// latency. The sequence of launches of a call is fully determined by its shapes and pointers, so the second time a
// call arrives with the same (shape, pointer) key it is stream-captured into a graph, and from then on the call is
// ONE cudaGraphLaunch. Per-call scalars that change (the Philox seed) live in device memory and are updated by a
// copy in front of the graph. PyTorch's caching allocator hands out the same blocks in a steady training loop, so
// the key (which includes the output / gradient pointers) repeats; when it does not, the call simply runs eagerly.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "prof.cuh"

namespace sscvae {

struct GraphEntry {
  std::vector<uint64_t> key;
  cudaGraphExec_t exec = nullptr;
  int state = 0;                       // 0: seen once (ran eagerly), 1: graph ready, -1: capture failed, stay eager
  uint64_t last_use = 0;
  unsigned long long launches = 0;     // kernels inside the graph (keeps sscvae_launch_count truthful)
};

struct GraphCache {
  std::vector<GraphEntry> entries;
  uint64_t tick = 0;
  cudaStream_t capture_stream = nullptr;   // the caller's stream may be the legacy default stream, which cannot capture
  ~GraphCache() {
    for (GraphEntry& e : entries)
      if (e.exec) cudaGraphExecDestroy(e.exec);
    if (capture_stream) cudaStreamDestroy(capture_stream);
  }
};

inline bool graphs_enabled() {
  static const bool off = [] { const char* e = getenv("SSCVAE_NO_GRAPHS"); return e && e[0] == '1'; }();
  return !off && !g_prof_enabled;
}

// body(stream) enqueues the whole call on `stream` and returns 0 / error.
template <class Body>
int run_with_graph(GraphCache& c, const std::vector<uint64_t>& key, cudaStream_t s, bool allow, Body&& body) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (!allow || !graphs_enabled() || cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return body(s);
  }
  ++c.tick;
  GraphEntry* e = nullptr;
  for (GraphEntry& x : c.entries)
    if (x.key == key) { e = &x; break; }
  if (!e) {
    if (c.entries.size() >= 8) {                       // evict the least recently used entry
      size_t lru = 0;
      for (size_t i = 1; i < c.entries.size(); ++i)
        if (c.entries[i].last_use < c.entries[lru].last_use) lru = i;
      if (c.entries[lru].exec) cudaGraphExecDestroy(c.entries[lru].exec);
      c.entries.erase(c.entries.begin() + lru);
    }
    c.entries.emplace_back();
    e = &c.entries.back();
    e->key = key;
    e->last_use = c.tick;
    return body(s);                                    // first sighting: eager (also runs every one-time initialisation)
  }
  e->last_use = c.tick;
  if (e->state == 1) {
    CUDA_TRY(cudaGraphLaunch(e->exec, s));
    g_launch_count += e->launches;
    return 0;
  }
  if (e->state < 0) return body(s);
  // second sighting: capture on the library's own stream, instantiate, launch on the caller's stream
  if (!c.capture_stream) CUDA_TRY(cudaStreamCreateWithFlags(&c.capture_stream, cudaStreamNonBlocking));
  const unsigned long long before_gemm = g_launch_count, before_pw = g_launch_count_pw;
  if (cudaStreamBeginCapture(c.capture_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    e->state = -1;
    return body(s);
  }
  const int rc = body(c.capture_stream);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(c.capture_stream, &graph);
  const unsigned long long captured = (g_launch_count - before_gemm) + (g_launch_count_pw - before_pw);
  g_launch_count = before_gemm;                         // nothing has run yet: the launches are counted per replay
  g_launch_count_pw = before_pw;
  if (rc != 0 || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    e->state = -1;
    return rc != 0 ? rc : body(s);
  }
  if (cudaGraphInstantiate(&e->exec, graph, 0) != cudaSuccess) {
    cudaGraphDestroy(graph);
    cudaGetLastError();
    e->exec = nullptr;
    e->state = -1;
    return body(s);
  }
  cudaGraphDestroy(graph);
  e->state = 1;
  e->launches = captured;
  CUDA_TRY(cudaGraphLaunch(e->exec, s));
  g_launch_count += e->launches;
  return 0;
}

inline void key_add(std::vector<uint64_t>& k, const void* p) { k.push_back(reinterpret_cast<uint64_t>(p)); }
inline void key_add(std::vector<uint64_t>& k, uint64_t v) { k.push_back(v); }

}  // namespace sscvae
