// C ABI: eval-mode decode (greedy / beam / constrained beam search) entirely on the device, plus the
// stand-alone search entry points used by the replay parity tests (include/sscvae.h).
//
// Rows r = (b*S + s)*K + k share image b's region features, projected features and mean-feature gate
// block through a row->image map; nothing image-sized is replicated per beam (the reference
// re-materialises (B*S*K, N, F) and re-projects it every step, updown_captioner.py:405-416).
#include "api_internal.cuh"
#include "search.cuh"

namespace sscvae {

// J > 1: "diverse sampling" (var_updown/scripts/inference.py:138-167 calls the model N_Z_SAMPLES times per image): every
// image is decoded J times with independent latent draws in ONE call. The search sees Bv = B*J independent sequences
// (virtual images); everything image-sized (features, projections, mean-feature gate block) stays at B and is shared
// through the row -> image map, so the per-step GEMMs run with M = B*J rows instead of J launches with M = B.
static void plan_decode(const Dims& d, int B, int N, int S, int K, int J, Plan& p) {
  const size_t b = sizeof(bf16), f = 4;
  const size_t Bv = (size_t)B * J, R = Bv * S * K, BN = (size_t)B * N;
  const int Pmax = K;
  p.add("seed", 16);
  p.add("featsb", BN * d.Fp * b);
  p.add("mask", BN * f);
  p.add("avgb", (size_t)B * d.Fp * b);
  p.add("projb", BN * d.Ap * b);
  p.add("gavg", (size_t)B * d.GP * f);
  p.add("pm_row", B * f);
  p.add("sent", B * f);
  p.add("rowmap", R * 4);                   // row -> image
  p.add("rowmap_exp", R * 4);               // row -> its sequence's step-0 row
  p.add("rowmap0", Bv * 4);                 // step-0 row -> image
  p.add("start_tok", Bv * 4);
  p.add("XA0", R * 2 * d.Hp * b);
  p.add("XA1", R * 2 * d.Hp * b);
  p.add("c1a", R * d.H * f);
  p.add("c1b", R * d.H * f);
  p.add("cda", R * d.H * f);
  p.add("cdb", R * d.H * f);
  p.add("XE", R * (d.Fp + d.Hp) * b);
  p.add("embb_r", R * d.Ep * b);
  p.add("ZB", R * d.ZC * b);                 // [z | c]: c = the conditioning block of sentiment_vae == 2
  if (d.cvar) p.add("pm", R * d.Z * f);      // per-row prior mean of the current step (updown_cell.py:160-163)
  p.add("acc", R * d.GP * f);
  p.add("q", R * d.A * f);
  p.add("alpha", R * N * f);
  if (d.tied) p.add("ob", R * d.Ep * b);
  p.add("logits", R * d.V * f);
  const size_t nst = gemm_rowstats_tiles(d.V);            // greedy: softmax statistics of the head GEMM instead of logits
  p.add("rs_max", R * nst * f);
  p.add("rs_sum", R * nst * f);
  p.add("rs_arg", R * nst * 4);
  p.add("fsm_bits", Bv * S * d.V * 4);
  p.add("cand_val", R * S * Pmax * f);
  p.add("cand_tok", R * S * Pmax * 4);
  p.add("tok_hist", (size_t)d.L * R * 4);
  p.add("bp_hist", (size_t)d.L * R * 4);
  p.add("score_hist", (size_t)d.L * R * f);
  p.add("best_samples", Bv * d.L * 8);                 // sscvae_decode_samples: the "best beam" copy of its predictions
}

const Plan& Handle::decode_plan(int B, int N, int S, int K, int J) {
  if (dp_B != B || dp_N != N || dp_S != S || dp_K != K || dp_J != J) {
    dp = Plan();
    plan_decode(d, B, N, S, K, J, dp);
    dp_B = B; dp_N = N; dp_S = S; dp_K = K; dp_J = J;
  }
  return dp;
}

static inline GemmSeg seg(const bf16* A, int lda, const bf16* B, int ldb, int K) {
  GemmSeg s; s.A = A; s.lda = lda; s.B = B; s.ldb = ldb; s.K = K; return s;
}

static int decode_impl(Handle* h, int B, int J, int N, int S, int K, int P, const char* pk, const void* const* wv,
                       const void* feats, const float* sent, const float* obj, const uint8_t* fsm, const long long* num_constraints,
                       int min_sat, const float* eps, unsigned long long seed, char* ws, size_t ws_bytes,
                       long long* predictions, float* log_probs, long long* best, int32_t* n_steps, cudaStream_t s) {
  const Dims& d = h->d;
  REQUIRE(B > 0 && N > 0 && S >= 1 && S <= 32 && K >= 1 && K <= 8 && P >= 1 && P <= K, "bad decode shape B=%d N=%d S=%d K=%d P=%d", B, N, S, K, P);
  const bool csent = d.cond && !d.cvar;
  REQUIRE(!csent || sent != nullptr, "sentiment is required when sentiment_vae == 1");
  REQUIRE(!d.cvar || obj != nullptr, "obj_means is required when sentiment_vae == 2");
  REQUIRE(S == 1 || fsm != nullptr, "an FSM is required for more than one state");
  REQUIRE(J >= 1 && (J == 1 || (S == 1 && fsm == nullptr)), "samples per image > 1 needs the unconstrained search (S == 1)");
  const Plan& dp = h->decode_plan(B, N, S, K, J);
  const int Bv = B * J;
  if (ws_bytes < dp.total) { set_error("workspace too small: %zu < %zu", ws_bytes, dp.total); return SSCVAE_ERR_WORKSPACE; }
  REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0 && (reinterpret_cast<uintptr_t>(pk) & 255) == 0,
          "workspace and packed weights must be 256-byte aligned");
  const Plan& pp = h->pp;
  auto W = [&](int i) { return reinterpret_cast<const float*>(wv[i]); };
  auto Pb = [&](const char* n) { return reinterpret_cast<const bf16*>(pk + pp.find(n)->off); };
  auto Pf = [&](const char* n) { return reinterpret_cast<const float*>(pk + pp.find(n)->off); };
  auto Wb = [&](const char* n) { return reinterpret_cast<bf16*>(ws + dp.find(n)->off); };
  auto Wf = [&](const char* n) { return reinterpret_cast<float*>(ws + dp.find(n)->off); };
  auto Wi = [&](const char* n) { return reinterpret_cast<int*>(ws + dp.find(n)->off); };
  auto zero = [&](const char* n) { return cudaMemsetAsync(ws + dp.find(n)->off, 0, dp.find(n)->bytes, s); };
  const int SK = S * K, R = Bv * SK, GP = d.GP, H = d.H, Hp = d.Hp, Fp = d.Fp, KXe = d.Fp + d.Hp, KX = d.KX, L = d.L;
  const unsigned long long* seed_dev = reinterpret_cast<const unsigned long long*>(ws + dp.find("seed")->off);
  (void)seed;

  CUDA_TRY(zero("XA0")); CUDA_TRY(zero("XA1")); CUDA_TRY(zero("XE")); CUDA_TRY(zero("bp_hist"));
  if (d.cvar) CUDA_TRY(zero("ZB"));                    // the padding columns behind c
  if (!h->opt_reuse_image_state) CUDA_TRY(zero("projb"));
  if (d.tied) CUDA_TRY(zero("ob"));
  // The per-image state (bf16 features, mask, mean, W_v projection, mean-feature gate block) depends on the images only:
  // a caller that decodes the SAME batch again (the reference's loop over latent samples, inference.py:138-167; its
  // lru_cache on the projected features, attention.py:99) sets the option and the state in the workspace is reused.
  const bool reuse = h->opt_reuse_image_state != 0;
  if (!reuse) TRY(image_prep(s, feats, h->opt_features_bf16, B, N, d.F, Wb("featsb"), Fp, Wf("mask"), Wb("avgb")));
  TRY(scale_rows_f32(s, csent ? sent : nullptr, d.mult, Wf("pm_row"), B));
  TRY(scale_rows_f32(s, csent ? sent : nullptr, 1.0f, Wf("sent"), B));
  TRY(iota_div_i32(s, Wi("rowmap"), R, SK * J));
  TRY(iota_div_i32(s, Wi("rowmap_exp"), R, SK));
  TRY(iota_div_i32(s, Wi("rowmap0"), Bv, J));
  TRY(fill_i32(s, Wi("start_tok"), d.boundary, Bv));                       // updown_captioner.py:326
  const uint32_t* fsm_bits = nullptr;
  if (fsm && h->opt_fsm_packed) {                      // already the (B,S,V) uint32 bit table (sscvae_fsm_build / sscvae_fsm_pack)
    fsm_bits = reinterpret_cast<const uint32_t*>(fsm);
  } else if (fsm) {
    TRY(fsm_pack(s, fsm, Bv, S, d.V, reinterpret_cast<uint32_t*>(Wi("fsm_bits"))));
    fsm_bits = reinterpret_cast<uint32_t*>(Wi("fsm_bits"));
  }
  if (!reuse) {  // once per image (the reference recomputes both every step in decode)
    GemmSeg sg = seg(Wb("featsb"), Fp, Pb("wv"), Fp, d.F);
    GemmEpi e; e.tag = "gemm.decode"; e.C16 = Wb("projb"); e.ldc16 = d.Ap;
    TRY(gemm_bf16_tn(s, B * N, d.A, 1, &sg, e));
    GemmSeg sa = seg(Wb("avgb"), Fp, Pb("w_att_f"), Fp, d.F);
    GemmEpi ea; ea.tag = "gemm.decode"; ea.C32 = Wf("gavg"); ea.ldc32 = GP;
    TRY(gemm_bf16_tn(s, B, GP, 1, &sa, ea));
  }
  int* tok_hist = Wi("tok_hist"); int* bp_hist = Wi("bp_hist"); float* score_hist = Wf("score_hist");
  bf16* XA[2] = {Wb("XA0"), Wb("XA1")};
  float* c1[2] = {Wf("c1a"), Wf("c1b")};
  float* cd[2] = {Wf("cda"), Wf("cdb")};

  // Unconstrained K = 1 search (greedy decode, diverse sampling): arg max and log-sum-exp come out of the head GEMM's
  // epilogue (RowStatsEpi) and the (rows, V) fp32 logits are never written: -2 x rows x V x 4 bytes per step.
  static const bool no_fused_head = [] { const char* e = getenv("SSCVAE_DECODE_FUSED_HEAD"); return e && e[0] == '0'; }();
  const bool greedy = S == 1 && K == 1 && fsm == nullptr && !no_fused_head;
  RowStatsEpi rs;
  rs.mode = 1; rs.st_max = Wf("rs_max"); rs.st_sum = Wf("rs_sum"); rs.st_arg = Wi("rs_arg");
  const int nst = gemm_rowstats_tiles(d.V);
  auto cell = [&](int rows, const int* tokens, const int* rowmap, bool first, const float* eps_t, int eps_stride,
                  int step) -> int {
    // in = index 0, out = index 1
    TRY(embed_gather_rows(s, tokens, rows, Pb("embb"), d.Ep, Wb("embb_r")));
    {
      LstmFwdArgs l = {};
      l.R = rows; l.H = H; l.add2 = Wf("gavg"); l.ld2 = GP; l.rowmap = rowmap; l.bias = Pf("b_att");
      l.c_prev = first ? nullptr : c1[0]; l.c_out = c1[1];
      l.h1_dst = Wb("XE") + Fp; l.ld_h1 = KXe; l.h2_dst = XA[1]; l.ld_h2 = 2 * Hp;
      GemmSeg sg[2] = {seg(Wb("embb_r"), d.Ep, Pb("w_att_e"), d.Ep, d.E), seg(XA[0], 2 * Hp, Pb("w_att_rec"), 2 * Hp, 2 * Hp)};
      GemmEpi e; e.tag = "gemm.decode"; e.C32 = Wf("acc"); e.ldc32 = GP; e.lstm = &l;   // cell fused when rows <= 256
      TRY(gemm_bf16_tn(s, rows, GP, first ? 1 : 2, sg, e));
    }
    {
      GemmSeg sg = seg(Wb("XE") + Fp, KXe, Pb("wq"), Hp, Hp);
      GemmEpi e; e.tag = "gemm.decode"; e.C32 = Wf("q"); e.ldc32 = d.A;
      TRY(gemm_bf16_tn(s, rows, d.A, 1, &sg, e));
      AttnArgs aa = {}; aa.R = rows; aa.N = N; aa.A = d.A; aa.Ap = d.Ap; aa.F = d.F; aa.Fp = Fp; aa.rowmap = rowmap;
      aa.q = Wf("q"); aa.ld_q = d.A; aa.proj = Wb("projb"); aa.feats = Wb("featsb"); aa.mask = Wf("mask");
      aa.w_a = W(SSCVAE_W_ATT_VEC);
      aa.rows_per_image = first ? J : SK * J;                  // rowmap0: r / J ; rowmap: r / (SK * J)
      TRY(attention_forward(s, aa, Wf("alpha"), nullptr, Wb("XE"), KXe));
    }
    {  // eval: z ~ N(prior_mean, prior_var) (updown_cell.py:200-208); no encoder LSTM
      LatentArgs la = {}; la.R = rows; la.Z = d.Z; la.Zp = d.Zp; la.sentiment_vae = d.sv; la.prior_var = d.prior_std * d.prior_std;
      la.prior_mean_row = Wf("pm_row"); la.rowmap = rowmap;
      if (d.cvar) {  // attribute-grounded prior: mean of this step from the row's attention weights; also the block c
        TRY(prior_mean_forward(s, Wf("alpha"), obj, rowmap, rows, N, d.Z, Wf("pm"), Wb("ZB") + d.Zp, d.ZC, d.cond));
        la.prior_mean_full = Wf("pm");
      }
      TRY(latent_forward_eval(s, la, eps_t, eps_stride, seed_dev, (unsigned long long)step, Wb("ZB"), d.ZC));
    }
    {
      GemmSeg sg[3] = {seg(Wb("XE"), KXe, Pb("w_dec_x"), KX, KXe),
                       seg(Wb("ZB"), d.ZC, Pb("w_dec_z"), d.ZC, d.ZC),
                       seg(XA[0] + Hp, 2 * Hp, Pb("w_dec_x") + KXe, KX, Hp)};
      LstmFwdArgs l = {};
      l.R = rows; l.H = H; l.bias = Pf("b_dec"); l.rowmap = rowmap;
      if (csent) { l.sent = Wf("sent"); l.scol = Pf("scol_dec"); }
      l.c_prev = first ? nullptr : cd[0]; l.c_out = cd[1];
      l.h1_dst = XA[1] + Hp; l.ld_h1 = 2 * Hp;
      GemmEpi e; e.tag = "gemm.decode"; e.C32 = Wf("acc"); e.ldc32 = GP; e.lstm = &l;
      TRY(gemm_bf16_tn(s, rows, GP, first ? 2 : 3, sg, e));
    }
    if (d.tied) {
      GemmSeg sg = seg(XA[1] + Hp, 2 * Hp, Pb("w_out"), Hp, Hp);
      GemmEpi e; e.tag = "gemm.decode"; e.bias = W(SSCVAE_W_OUT_PROJ_B); e.act = 1; e.C16 = Wb("ob"); e.ldc16 = d.Ep;
      TRY(gemm_bf16_tn(s, rows, d.E, 1, &sg, e));
      GemmSeg sv = seg(Wb("ob"), d.Ep, Pb("embb"), d.Ep, d.E);
      GemmEpi ev; ev.tag = "gemm.decode";
      if (greedy) ev.rs = &rs; else { ev.C32 = Wf("logits"); ev.ldc32 = d.V; }
      TRY(gemm_bf16_tn(s, rows, d.V, 1, &sv, ev));
    } else {
      GemmSeg sg = seg(XA[1] + Hp, 2 * Hp, Pb("w_out"), Hp, Hp);
      GemmEpi e; e.tag = "gemm.decode"; e.bias = W(SSCVAE_W_OUT_PROJ_B);
      if (greedy) e.rs = &rs; else { e.C32 = Wf("logits"); e.ldc32 = d.V; }
      TRY(gemm_bf16_tn(s, rows, d.V, 1, &sg, e));
    }
    return 0;
  };

  // ---- step 0: one row per image, zero states (cbs.py:127-155)
  TRY(cell(Bv, Wi("start_tok"), Wi("rowmap0"), true, eps, SK, 0));
  if (greedy) {
    TRY(greedy_merge(s, rs.st_max, rs.st_sum, rs.st_arg, nst, Bv, nullptr, nullptr, d.boundary, tok_hist, nullptr, score_hist));
  } else {
    SearchRowsArgs a = {};
    a.logp = Wf("logits"); a.ld = d.V; a.V = d.V; a.normalized = 0; a.fsm_bits = fsm_bits; a.R = Bv; a.S = S; a.K = K;
    a.rows_per_image = 1; a.P = K; a.end_index = d.boundary; a.neg_value = -INFINITY;
    a.cand_val = score_hist; a.cand_tok = tok_hist;
    TRY(search_rows(s, a));
  }
  TRY(gather_rows_bf16(s, XA[1], Wi("rowmap_exp"), R, 2 * Hp, 2 * Hp, XA[0]));
  TRY(gather_rows_f32(s, c1[1], Wi("rowmap_exp"), R, H, c1[0]));
  TRY(gather_rows_f32(s, cd[1], Wi("rowmap_exp"), R, H, cd[0]));
  // ---- steps 1..L-1 on all R rows (cbs.py:161-250); the early exit is resolved at the end
  for (int t = 1; t < L; ++t) {
    TRY(cell(R, tok_hist + (size_t)(t - 1) * R, Wi("rowmap"), false, eps ? eps + (size_t)t * R * d.Z : nullptr, 1, t));
    if (greedy) {
      TRY(greedy_merge(s, rs.st_max, rs.st_sum, rs.st_arg, nst, R, tok_hist + (size_t)(t - 1) * R, score_hist + (size_t)(t - 1) * R,
                       d.boundary, tok_hist + (size_t)t * R, bp_hist + (size_t)t * R, score_hist + (size_t)t * R));
      TRY(state_gather(s, bp_hist + (size_t)t * R, R, SK, XA[1], XA[0], 2 * Hp, c1[1], c1[0], cd[1], cd[0], H));
      continue;
    }
    SearchRowsArgs a = {};
    a.logp = Wf("logits"); a.ld = d.V; a.V = d.V; a.normalized = 0; a.fsm_bits = fsm_bits; a.R = R; a.S = S; a.K = K;
    a.rows_per_image = SK; a.P = P; a.end_index = d.boundary; a.neg_value = -1e20f;
    a.last_tokens = tok_hist + (size_t)(t - 1) * R; a.last_scores = score_hist + (size_t)(t - 1) * R;
    a.cand_val = Wf("cand_val"); a.cand_tok = Wi("cand_tok");
    TRY(search_rows(s, a));
    TRY(search_merge(s, Wf("cand_val"), Wi("cand_tok"), Bv, S, K, P, tok_hist + (size_t)t * R, bp_hist + (size_t)t * R,
                     score_hist + (size_t)t * R));
    TRY(state_gather(s, bp_hist + (size_t)t * R, R, SK, XA[1], XA[0], 2 * Hp, c1[1], c1[0], cd[1], cd[0], H));
  }
  TRY(search_finish(s, tok_hist, bp_hist, score_hist, L, Bv, S, K, d.boundary, num_constraints, min_sat, predictions,
                    log_probs, best, n_steps));
  return 0;
}

}  // namespace sscvae

using namespace sscvae;

extern "C" {

int sscvae_fsm_pack(const uint8_t* fsm, int batch, int states, int vocab, uint32_t* fsm_bits, void* stream) {
  REQUIRE(fsm && fsm_bits && batch > 0 && vocab > 0, "bad argument");
  return fsm_pack(reinterpret_cast<cudaStream_t>(stream), fsm, batch, states, vocab, fsm_bits);
}

int sscvae_fsm_build(const int32_t* connections, const int32_t* connection_offsets, const int32_t* wordform_ids,
                     const int32_t* state_counts, int batch, int states, int vocab, uint32_t* fsm_bits, void* stream) {
  REQUIRE(connections && connection_offsets && wordform_ids && state_counts && fsm_bits && batch > 0 && vocab > 0, "bad argument");
  return fsm_build(reinterpret_cast<cudaStream_t>(stream), connections, connection_offsets, wordform_ids, state_counts, batch,
                   states, vocab, fsm_bits);
}

int sscvae_select_best_beam(const int64_t* predictions, const float* log_probs, const uint8_t* valid_states, int batch,
                            int states, int beam, int steps, int64_t* best, void* stream) {
  REQUIRE(predictions && log_probs && valid_states && best && batch > 0 && states > 0 && beam > 0 && steps > 0, "bad argument");
  return select_best_masked(reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const long long*>(predictions), log_probs,
                            valid_states, batch, states, beam, steps, reinterpret_cast<long long*>(best));
}

int sscvae_search_first_step(const float* logp, int batch, int states, int beam, int vocab, const uint32_t* fsm_bits,
                             int normalized, int32_t* tokens, float* scores, void* stream) {
  REQUIRE(logp && tokens && scores, "NULL argument");
  SearchRowsArgs a = {};
  a.logp = logp; a.ld = vocab; a.V = vocab; a.normalized = normalized; a.fsm_bits = fsm_bits; a.R = batch; a.S = states;
  a.K = beam; a.rows_per_image = 1; a.P = beam; a.end_index = -1; a.neg_value = -INFINITY;
  a.cand_val = scores; a.cand_tok = tokens;
  return search_rows(reinterpret_cast<cudaStream_t>(stream), a);
}

size_t sscvae_search_scratch_bytes(int batch, int states, int beam, int per_node) {
  return (size_t)batch * states * beam * states * per_node * 8 + 1024;
}

int sscvae_search_step(const float* logp, int batch, int states, int beam, int per_node, int vocab,
                       const uint32_t* fsm_bits, int normalized, int end_index, const int32_t* last_tokens,
                       const float* last_scores, void* scratch, size_t scratch_bytes, int32_t* tokens, int32_t* backptr,
                       float* scores, void* stream) {
  REQUIRE(logp && last_tokens && last_scores && scratch && tokens && backptr && scores, "NULL argument");
  const int R = batch * states * beam;
  const size_t n = (size_t)R * states * per_node;
  if (scratch_bytes < n * 8) { set_error("search scratch too small"); return SSCVAE_ERR_WORKSPACE; }
  float* cand_val = reinterpret_cast<float*>(scratch);
  int32_t* cand_tok = reinterpret_cast<int32_t*>(cand_val + n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SearchRowsArgs a = {};
  a.logp = logp; a.ld = vocab; a.V = vocab; a.normalized = normalized; a.fsm_bits = fsm_bits; a.R = R; a.S = states;
  a.K = beam; a.rows_per_image = states * beam; a.P = per_node; a.end_index = end_index; a.neg_value = -1e20f;
  a.last_tokens = last_tokens; a.last_scores = last_scores; a.cand_val = cand_val; a.cand_tok = cand_tok;
  TRY(search_rows(st, a));
  return search_merge(st, cand_val, cand_tok, batch, states, beam, per_node, tokens, backptr, scores);
}

int sscvae_search_finish(const int32_t* tokens_hist, const int32_t* backptr_hist, const float* scores_hist, int steps_run,
                         int batch, int states, int beam, int end_index, const int64_t* num_constraints,
                         int min_constraints_to_satisfy, int64_t* predictions, float* final_scores, int64_t* best,
                         int32_t* n_steps, void* stream) {
  REQUIRE(tokens_hist && backptr_hist && scores_hist && predictions && final_scores && best && n_steps, "NULL argument");
  return search_finish(reinterpret_cast<cudaStream_t>(stream), tokens_hist, backptr_hist, scores_hist, steps_run, batch,
                       states, beam, end_index, reinterpret_cast<const long long*>(num_constraints),
                       min_constraints_to_satisfy, reinterpret_cast<long long*>(predictions), final_scores,
                       reinterpret_cast<long long*>(best), n_steps);
}

size_t sscvae_decode_workspace_bytes(const SscvaeHandle* hh, int batch, int num_boxes, int states, int beam) {
  Handle* h = const_cast<Handle*>(reinterpret_cast<const Handle*>(hh));
  if (!h || batch <= 0 || num_boxes <= 0 || states <= 0 || beam <= 0) return 0;
  return h->decode_plan(batch, num_boxes, states, beam, 1).total;
}

size_t sscvae_decode_samples_workspace_bytes(const SscvaeHandle* hh, int batch, int samples, int num_boxes) {
  Handle* h = const_cast<Handle*>(reinterpret_cast<const Handle*>(hh));
  if (!h || batch <= 0 || num_boxes <= 0 || samples <= 0) return 0;
  return h->decode_plan(batch, num_boxes, 1, 1, samples).total;
}

int sscvae_decode_region(const SscvaeHandle* hh, int batch, int num_boxes, int states, int beam, const char* name,
                         size_t* offset, size_t* bytes) {
  Handle* h = const_cast<Handle*>(reinterpret_cast<const Handle*>(hh));
  REQUIRE(h && name && offset && bytes, "NULL argument");
  const Region* r = h->decode_plan(batch, num_boxes, states, beam, 1).find(name);
  REQUIRE(r != nullptr, "unknown workspace region '%s'", name);
  *offset = r->off; *bytes = r->bytes;
  return 0;
}

int sscvae_decode(SscvaeHandle* hh, int batch, int num_boxes, int states, int beam, int per_node, const void* packed,
                  const void* const* weights, const float* image_features, const float* sentiment, const float* obj_means,
                  const uint8_t* fsm, const int64_t* num_constraints, int min_constraints_to_satisfy, const float* eps, uint64_t seed,
                  void* workspace, size_t workspace_bytes, int64_t* predictions, float* log_probs, int64_t* best,
                  int32_t* n_steps, void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  REQUIRE(h && packed && weights && image_features && workspace && predictions && log_probs && best && n_steps, "NULL argument");
  REQUIRE(batch > 0 && num_boxes > 0 && states >= 1 && beam >= 1, "bad decode shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const Plan& dp = h->decode_plan(batch, num_boxes, states, beam, 1);
  if (workspace_bytes < dp.total) { set_error("workspace too small: %zu < %zu", workspace_bytes, dp.total); return SSCVAE_ERR_WORKSPACE; }
  const unsigned long long seed_host = seed;
  CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(workspace) + dp.find("seed")->off, &seed_host, sizeof(seed_host),
                           cudaMemcpyHostToDevice, st));
  std::vector<uint64_t> key;
  for (uint64_t v : {(uint64_t)batch, (uint64_t)num_boxes, (uint64_t)states, (uint64_t)beam, (uint64_t)per_node,
                     (uint64_t)min_constraints_to_satisfy, (uint64_t)workspace_bytes,
                     (uint64_t)(h->opt_fsm_packed * 4 + h->opt_features_bf16 * 2 + h->opt_reuse_image_state)})
    key_add(key, v);
  for (const void* q : {packed, (const void*)image_features, (const void*)sentiment, (const void*)obj_means, (const void*)fsm,
                        (const void*)num_constraints, (const void*)eps, (const void*)workspace, (const void*)predictions,
                        (const void*)log_probs, (const void*)best, (const void*)n_steps})
    key_add(key, q);
  for (int i = 0; i < SSCVAE_W_COUNT; ++i) key_add(key, weights[i]);
  return run_with_graph(h->dec_graphs, key, st, true, [&](cudaStream_t s) {
    return decode_impl(h, batch, 1, num_boxes, states, beam, per_node, reinterpret_cast<const char*>(packed), weights,
                       image_features, sentiment, obj_means, fsm, reinterpret_cast<const long long*>(num_constraints),
                       min_constraints_to_satisfy, eps, seed, reinterpret_cast<char*>(workspace), workspace_bytes,
                       reinterpret_cast<long long*>(predictions), log_probs, reinterpret_cast<long long*>(best), n_steps, s);
  });
}

int sscvae_decode_samples(SscvaeHandle* hh, int batch, int samples, int num_boxes, const void* packed,
                          const void* const* weights, const float* image_features, const float* sentiment,
                          const float* obj_means, const float* eps, uint64_t seed, void* workspace, size_t workspace_bytes, int64_t* predictions, float* log_probs,
                          int32_t* n_steps, void* stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  REQUIRE(h && packed && weights && image_features && workspace && predictions && log_probs && n_steps, "NULL argument");
  REQUIRE(batch > 0 && num_boxes > 0 && samples >= 1, "bad decode shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const Plan& dp = h->decode_plan(batch, num_boxes, 1, 1, samples);
  if (workspace_bytes < dp.total) { set_error("workspace too small: %zu < %zu", workspace_bytes, dp.total); return SSCVAE_ERR_WORKSPACE; }
  const unsigned long long seed_host = seed;
  CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(workspace) + dp.find("seed")->off, &seed_host, sizeof(seed_host),
                           cudaMemcpyHostToDevice, st));
  std::vector<uint64_t> key;
  for (uint64_t v : {(uint64_t)batch, (uint64_t)num_boxes, (uint64_t)samples, (uint64_t)0x5a5a, (uint64_t)workspace_bytes,
                     (uint64_t)(h->opt_features_bf16 * 2 + h->opt_reuse_image_state)})
    key_add(key, v);
  for (const void* q : {packed, (const void*)image_features, (const void*)sentiment, (const void*)obj_means, (const void*)eps,
                        (const void*)workspace, (const void*)predictions, (const void*)log_probs, (const void*)n_steps})
    key_add(key, q);
  for (int i = 0; i < SSCVAE_W_COUNT; ++i) key_add(key, weights[i]);
  char* wsb = reinterpret_cast<char*>(workspace);
  long long* best = reinterpret_cast<long long*>(wsb + dp.find("best_samples")->off);
  return run_with_graph(h->dec_graphs, key, st, true, [&](cudaStream_t s) {
    return decode_impl(h, batch, samples, num_boxes, 1, 1, 1, reinterpret_cast<const char*>(packed), weights, image_features,
                       sentiment, obj_means, nullptr, nullptr, 0, eps, seed, wsb, workspace_bytes, reinterpret_cast<long long*>(predictions),
                       log_probs, best, n_steps, s);
  });
}

}  // extern "C"
