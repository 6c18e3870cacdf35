// Internal declarations shared by the C-ABI translation units.
#pragma once
#include <algorithm>
#include <cstring>
#include <initializer_list>
#include <vector>
#include "../../include/sscvae.h"
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "graph_cache.cuh"

namespace sscvae {

constexpr int kPad = 64;   // elements: operand row strides / column blocks are multiples of 128 bytes (see init_dims)


struct Dims {
  int F, E, H, A, Z, V, L, T;
  int sv, simple, tied, pad, boundary, cond;
  // sentiment_vae == 2 (updown_cell.py:160-190): the conditioning block of the encoder / decoder LSTM inputs is the per-step
  // prior mean (cond = Z columns, or 1 with latent_embedding "senti_word_net"), a time-varying bf16 operand kept next to z:
  // rows of ZB are [z (Zp) | c (Cp)], ZC = Zp + Cp. cvar = 0: Cp = 0 and ZC = Zp.
  int le, cvar, Cp, ZC;
  float prior_std, mult;
  int Fp, Ep, Hp, Ap, Zp, Vp, G, Gp, Z2, Z2p, KX;
  int debug_logits;                    // SSCVAE_DEBUG_LOGITS=1 at create: the training forward also stores the fp32 logits (tests)
  int GP;                              // rows of a packed forward LSTM weight block: gate-interleaved, lstm_gate_rows(H)
};
int init_dims(const SscvaeDims* in, Dims& d);
int set_l2_window(cudaStream_t s, const void* base, size_t bytes);

struct Region { const char* name; size_t off, bytes; };
struct Plan {
  std::vector<Region> regs;
  size_t total = 0;
  void add(const char* name, size_t bytes);
  const Region* find(const char* name) const;
};

// Fork / join of independent launches over a few auxiliary streams (works inside a stream capture: the auxiliary streams
// join the capture through the fork event and are joined back before the call returns). Used for the weight-gradient
// GEMMs of one LSTM, which share an operand but are otherwise independent: launched back to back on one stream each
// of them ends in a partial wave on the 74 CTA pairs.
struct StreamFork {
  static constexpr int kAux = 3;
  cudaStream_t aux[kAux] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kAux] = {nullptr, nullptr, nullptr};
  bool forked = false;
  int init() {
    if (ev_fork) return 0;
    for (int i = 0; i < kAux; ++i) {
      CUDA_TRY(cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    return 0;
  }
  int fork(cudaStream_t s) {
    TRY(init());
    CUDA_TRY(cudaEventRecord(ev_fork, s));
    for (int i = 0; i < kAux; ++i) CUDA_TRY(cudaStreamWaitEvent(aux[i], ev_fork, 0));
    forked = true;
    return 0;
  }
  // stream of the i-th independent item
  cudaStream_t pick(cudaStream_t s, int i) const { return (!forked || i % (kAux + 1) == 0) ? s : aux[i % (kAux + 1) - 1]; }
  int join(cudaStream_t s) {
    if (!forked) return 0;
    for (int i = 0; i < kAux; ++i) {
      CUDA_TRY(cudaEventRecord(ev_join[i], aux[i]));
      CUDA_TRY(cudaStreamWaitEvent(s, ev_join[i], 0));
    }
    forked = false;
    return 0;
  }
  ~StreamFork() {
    for (int i = 0; i < kAux; ++i) {
      if (aux[i]) cudaStreamDestroy(aux[i]);
      if (ev_join[i]) cudaEventDestroy(ev_join[i]);
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
  }
};

struct Handle {
  Dims d;
  StreamFork fork;
  Plan pp;                 // packed weights
  Plan tp; int tp_B = -1, tp_N = -1;   // training workspace for the last (B,N)
  Plan dp; int dp_B = -1, dp_N = -1, dp_S = -1, dp_K = -1, dp_J = -1;   // decode workspace
  GraphCache fwd_graphs, bwd_graphs, dec_graphs;
  // per-handle options (sscvae_set_option), consumed by the following calls
  int opt_features_bf16 = 0;           // image_features pointers are bf16 (B,N,F) instead of fp32
  int opt_reuse_image_state = 0;       // decode: the workspace already holds this batch's featsb / projb / mask / avg state
  int opt_persistent_bwd = 1;          // training: run the BPTT loop as the persistent kernel (recurrent_bwd.cu) when the shape fits
  int opt_fsm_packed = 0;              // decode: `fsm` is the (B,S,V) uint32 bit table, not the (B,S,S,V) uint8 tensor
  const Plan& train_plan(int B, int N);
  const Plan& decode_plan(int B, int N, int S, int K, int J);
};

}  // namespace sscvae
