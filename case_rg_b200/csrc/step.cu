// Error plumbing and the whole-step orchestrators: one C call enqueues every kernel of a decode
// step on the caller's stream (so a step, or a whole decode, can be captured in a CUDA graph).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cb {

static thread_local char g_err[512] = "ok";
int g_use_pdl = 1;
int g_use_chain = 1;
int g_use_fork = 1;
int g_use_stack = 1;
int g_use_tail = 2;
int g_use_fused_select = 1;
int g_use_post = 1;
int g_use_gate = 1;
int g_use_gate_h = 0;          // f16 / tensor-core form of the gate kernel when the step arguments carry U16 (measured slower: off)
int g_evict_first = 1;         // L2 evict-first policy on the once-per-step K|V and Uk.mem streams
int g_use_plan = 1;            // sparse tail from the prefill's copy plan (sorted unique ids) instead of the hash table
int g_xnext = 0;               // tiles per warp the passage cross-attention prefetches for the next layer's launch
int g_prefetch_pct = 0;        // measured: no gain at C2 (the prefetch traffic slows the latency-bound launch more than it helps)
int g_prefetch_mask = 3;       // bit 0: from the stack launch (layer 4), bit 1: from the chain launches (layers 5..7)
// side stream + events for the fork/join inside a step (created on first use, outside any capture: the
// engines run one uncaptured warm-up step before they capture)
static cudaStream_t g_aux = nullptr;
static cudaEvent_t g_ev_fork[2] = {nullptr, nullptr}, g_ev_join[2] = {nullptr, nullptr};
static bool aux_ready() {
  if (g_aux) return true;
  if (cudaStreamCreateWithFlags(&g_aux, cudaStreamNonBlocking) != cudaSuccess) { g_aux = nullptr; return false; }
  for (int i = 0; i < 2; ++i) {
    if (cudaEventCreateWithFlags(&g_ev_fork[i], cudaEventDisableTiming) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&g_ev_join[i], cudaEventDisableTiming) != cudaSuccess) return false;
  }
  return true;
}
cudaError_t g_launch_err = cudaSuccess;

void set_error(const char* msg) {
  strncpy(g_err, msg, sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = g_launch_err;
  g_launch_err = cudaSuccess;
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace cb

namespace cb {
// L2 warm-up for the NEXT step's first launch (the fused query-memory stack): what it reads - 4 MB of layer weights,
// the query memory's K|V tiles, the self-attention history of its four layers - was pushed out of L2 by the step's
// 0.6 GB of K|V streams (experiment, default off: measured no gain, the launch's 77 us in the graph against 54-60 us
// stand-alone are not L2 misses on these regions).  Launched on the side stream behind the passage additive attention.
struct PfRegions { const char* p[16]; unsigned long long bytes[16]; int n; };
__global__ void l2_prefetch_kernel(PfRegions r) {
  pdl_wait();
  constexpr unsigned CH = 4096;
  for (int i = 0; i < r.n; ++i) {
    const unsigned long long nch = (r.bytes[i] + CH - 1) / CH;
    for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < nch;
         c += (unsigned long long)gridDim.x * blockDim.x) {
      const unsigned long long off = c * CH;
      const unsigned len = (unsigned)min((unsigned long long)CH, r.bytes[i] - off) & ~15u;
      if (len) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(r.p[i] + off), "r"(len) : "memory");
    }
  }
}
}  // namespace cb
int g_next_prefetch = 0;       // bit 0: weights + query-memory K|V, bit 1: self-attention history of layers 0..3.  Measured
                               // at C2: the stack launch stays at 77 us and the step gets slower (0.333 -> 0.345 / 0.404 ms): off

using namespace cb;

extern "C" int case_set_next_step_prefetch(int mask) {
  const int old = g_next_prefetch;
  if (mask >= 0) g_next_prefetch = mask;
  return old;
}

extern "C" int case_abi_version(void) { return 1; }
extern "C" int case_set_pdl(int on) {
  const int old = g_use_pdl;
  g_use_pdl = on ? 1 : 0;
  return old;
}
extern "C" int case_set_chain(int on) {
  const int old = g_use_chain;
  g_use_chain = on ? 1 : 0;
  return old;
}
extern "C" int case_set_fused_tail(int on) {
  const int old = g_use_tail;
  g_use_tail = on < 0 ? 0 : (on > 2 ? 2 : on);
  return old;
}
extern "C" int case_set_stack_fusion(int on) {
  const int old = g_use_stack;
  g_use_stack = on ? 1 : 0;
  return old;
}
extern "C" int case_set_fused_select(int on) {
  const int old = g_use_fused_select;
  g_use_fused_select = on ? 1 : 0;
  return old;
}
extern "C" int case_set_post_linears(int on) {
  const int old = g_use_post;
  g_use_post = on ? 1 : 0;
  return old;
}
extern "C" int case_set_gate_form(int on) {
  const int old = g_use_gate;
  if (on >= 0) g_use_gate = on ? 1 : 0;      // negative: query only
  return old;
}
extern "C" int case_set_stream_evict_first(int on) {
  const int old = g_evict_first;
  if (on >= 0) g_evict_first = on ? 1 : 0;   // negative: query only
  return old;
}
extern "C" int case_set_copy_plan(int on) {
  const int old = g_use_plan;
  if (on >= 0) g_use_plan = on ? 1 : 0;
  return old;
}
extern "C" int case_set_gate_f16(int on) {
  const int old = g_use_gate_h;
  if (on >= 0) g_use_gate_h = on ? 1 : 0;
  return old;
}
extern "C" int case_set_xattn_next_prefetch(int ntiles) {
  const int old = g_xnext;
  g_xnext = ntiles < 0 ? 0 : ntiles;
  return old;
}
extern "C" int case_set_kv_prefetch(int pct) {
  const int old = g_prefetch_pct;
  g_prefetch_pct = pct < 0 ? 0 : (pct > 100 ? 100 : pct);
  if (getenv("CASE_PF_MASK")) g_prefetch_mask = atoi(getenv("CASE_PF_MASK"));
  return old;
}
extern "C" int case_set_fork(int on) {
  const int old = g_use_fork;
  g_use_fork = on ? 1 : 0;
  return old;
}
extern "C" const char* case_last_error(void) { return g_err; }

#define TRY(x)            \
  do {                    \
    int _e = (x);         \
    if (_e) return _e;    \
  } while (0)

static case_seg_t seg(const float* p, int ld, int width, int div, int gather = 0) {
  case_seg_t s;
  s.p = p; s.ld = ld; s.width = width; s.div = div; s.gather = gather;
  return s;
}

static case_select_args_t select_args(int mode, int B, int W, int t, int max_len, int Tmax, int BOS, int EOS, int UNK,
                                      int PAD, const float* top_vals, const int32_t* top_idx, int32_t* live, double* cum,
                                      int32_t* length, int32_t* tok, int32_t* const anc[2], int32_t* parent,
                                      int32_t* ended, double* best_key, int32_t* best_len, int32_t* out_tokens,
                                      int32_t* n_live) {
  case_select_args_t s;
  memset(&s, 0, sizeof(s));
  s.mode = mode; s.B = B; s.W = W; s.t = t; s.max_len = max_len; s.Tmax = Tmax;
  s.BOS = BOS; s.EOS = EOS; s.UNK = UNK; s.PAD = PAD;
  s.top_vals = top_vals; s.top_idx = top_idx; s.live = live; s.cum = cum; s.length = length; s.tok = tok;
  s.anc_in = anc[t & 1]; s.anc_out = anc[(t + 1) & 1]; s.parent = parent; s.ended = ended;
  s.best_key = best_key; s.best_len = best_len; s.out_tokens = out_tokens; s.n_live = n_live;
  return s;
}

static int select_step(int mode, int B, int W, int t, int max_len, int Tmax, int BOS, int EOS, int UNK, int PAD,
                       const float* top_vals, const int32_t* top_idx, int32_t* live, double* cum, int32_t* length,
                       int32_t* tok, int32_t* const anc[2], int32_t* parent, int32_t* ended, double* best_key,
                       int32_t* best_len, int32_t* out_tokens, int32_t* n_live, cudaStream_t st) {
  case_select_args_t s;
  memset(&s, 0, sizeof(s));
  s.mode = mode; s.B = B; s.W = W; s.t = t; s.max_len = max_len; s.Tmax = Tmax;
  s.BOS = BOS; s.EOS = EOS; s.UNK = UNK; s.PAD = PAD;
  s.top_vals = top_vals; s.top_idx = top_idx; s.live = live; s.cum = cum; s.length = length; s.tok = tok;
  s.anc_in = anc[t & 1]; s.anc_out = anc[(t + 1) & 1]; s.parent = parent; s.ended = ended;
  s.best_key = best_key; s.best_len = best_len; s.out_tokens = out_tokens; s.n_live = n_live;
  return case_beam_select(&s, st);
}

extern "C" int case_decode_step(const case_step_args_t* a, int t, case_stream_t stream) {
  CB_REQUIRE(a, "case_decode_step: null args");
  CB_REQUIRE(a->R == a->B * a->W && a->W >= 1 && a->W <= CASE_MAX_W, "case_decode_step: R != B*W or W out of range");
  CB_REQUIRE(t >= 0 && t < a->Tmax && a->Tmax <= CASE_MAX_T, "case_decode_step: t out of range");
  CB_REQUIRE(a->materialize_only || t < a->max_len, "case_decode_step: t >= max_len");
  cudaStream_t st = (cudaStream_t)stream;
  const int R = a->R, B = a->B, W = a->W, TL = a->Tmax + 1, dt = a->dtype;
  const int32_t* anc = a->anc[t & 1];

  // search path: the sparse tail (touched ids + base candidates); `generate` face: the dense fused tail
  const int k2 = 2 * W;
  const bool sparse = g_use_tail == 2 && !a->materialize_only && a->base_ms && a->base_e && a->base_i && a->V >= k2 &&
                      a->S[0] + a->S[1] <= case_sparse_tail_max_sources();
  // gate form of the additive attentions: the contexts only feed the mixture gate, so 3 gate-projected numbers
  // per key replace the value rows (needs the sparse tail, which merges gate partials instead of contexts)
  const bool gate = g_use_gate && sparse && dt == CASE_BF16 && a->Gv[0] != nullptr && a->Gv[1] != nullptr;
  // attns[i]: query = [dec_out ; norm2(answer_rep)]   (Model.py:108), then the fused additive attention
  auto stack_attention = [&](int i, const float* hsrc, float* qa, cudaStream_t s2, bool have_qa = false) -> int {
    if (have_qa) goto additive;          // the query was produced by a post linear of the preceding cluster launch
    {
    case_rowlin_args_t q;
    memset(&q, 0, sizeof(q));
    q.seg[0] = seg(hsrc, H, H, 1);
    q.seg[1] = seg(a->feat, H, H, W);
    q.nseg = 2; q.K = 2 * H; q.Wt = a->Wqa_t[i]; q.bias = a->bqa[i]; q.N = H; q.out = qa; q.ldo = H;
    q.R = R; q.dtype = dt;
    TRY(case_row_linear(&q, s2));
    }
  additive:
    if (gate) {
      const bool cmp = i == 1 && a->xidx != nullptr && a->xcount != nullptr;
      if (g_use_gate_h && a->U16[i] != nullptr && W >= 2 && a->fast_tanh)
        return case_additive_attn_gate_h(qa, a->U16[i], a->Gv[i], a->va[i], a->mask[i], a->prior[i], a->tok, TL, t, B, W,
                                         a->S[i], a->nsplit_a[i], a->attn_un[i], a->stats[i], a->ctxp[i],
                                         cmp ? a->xidx : nullptr, cmp ? a->xcount : nullptr, cmp ? a->xorder : nullptr,
                                         cmp ? a->xns : nullptr, s2);
      return case_additive_attn_gate(qa, a->U[i], a->Gv[i], a->va[i], a->mask[i], a->prior[i], a->tok, TL, t, B, W, a->S[i],
                                     a->nsplit_a[i], a->attn_un[i], a->stats[i], a->ctxp[i], a->fast_tanh,
                                     cmp ? a->xidx : nullptr, cmp ? a->xcount : nullptr, cmp ? a->xorder : nullptr,
                                     cmp ? a->xns : nullptr, s2);
    }
    if (i == 1 && dt == CASE_BF16 && a->xidx != nullptr && a->xcount != nullptr)   // valid keys only, balanced splits
      return case_additive_attn_compact(qa, a->U[i], a->Mv[i], a->va[i], a->mask[i], a->prior[i], a->tok, TL, t, B, W,
                                        a->S[i], H, a->nsplit_a[i], a->attn_un[i], a->stats[i], a->ctxp[i],
                                        a->fast_tanh, a->xidx, a->xcount, a->xorder, s2);
    return case_additive_attn(qa, a->U[i], a->Mv[i], a->va[i], a->mask[i], a->prior[i], a->tok, TL, t, B, W, a->S[i],
                              H, a->nsplit_a[i], a->attn_un[i], a->stats[i], a->ctxp[i], a->fast_tanh, dt, s2);
  };
  auto gen0 = [&]() -> int {   // gen.0 on [dec_input ; norm1(dec_out) ; feat]   (Model.py:115)
    case_rowlin_args_t g;
    memset(&g, 0, sizeof(g));
    g.seg[0] = seg(a->x_in, H, H, 1);
    g.seg[1] = seg(a->hN, H, H, 1);
    g.seg[2] = seg(a->feat, H, H, W);
    g.nseg = 3; g.K = 3 * H; g.Wt = a->Wg_t; g.bias = a->bg; g.N = H; g.out = a->gfeat; g.ldo = H;
    g.R = R; g.dtype = dt;
    return case_row_linear(&g, st);
  };
  auto finalize = [&]() -> int {
    return case_finalize_rows(a->h, a->lnN_g, a->lnN_b, a->stats[0], a->ctxp[0], a->nsplit_a[0], a->stats[1],
                              a->ctxp[1], a->nsplit_a[1], a->Wm, a->bm, a->hN, a->ctx[0], a->ctx[1], a->gates, a->fac, R,
                              st);
  };
#define CUTRY(x)                                                         \
  do {                                                                   \
    cudaError_t _ce = (x);                                               \
    if (_ce != cudaSuccess) {                                            \
      set_error(cudaGetErrorString(_ce));                                \
      return (int)_ce;                                                   \
    }                                                                    \
  } while (0)
  // cross-attention over memory i for layer L: compacted + balanced partition where the prefill provided it
  const bool xpart = dt == CASE_BF16 && a->xcount != nullptr && a->xprefix != nullptr && a->xslots > 0;
  auto big_xattn = [&](int L) -> int {
    const int i = L / 4;
    if (i == 1 && xpart && g_xnext > 0 && L < 7) case_cross_attn_part_next(a->Kx[L + 1], g_xnext);
    if (i == 1 && xpart)
      return case_cross_attn_part(a->q2, a->Kx[L], a->xcount, a->xprefix, B, W, a->S[1], a->xslots, a->part_ml, a->part_acc,
                                  st);
    return case_cross_attn_partial_tc(a->q2, a->Kx[L], a->mask[i], B, W, a->S[i], a->nsplit_x[i], a->part_ml,
                                      a->part_acc, st);
  };
  auto nparts_of = [&](int L) -> int { return (L / 4 == 1 && xpart) ? a->xslots : a->nsplit_x[L / 4]; };
  const bool chain = dt == CASE_BF16 && g_use_chain && a->layers[0].Wc != nullptr && a->Tmax <= case_layer_chain_max_tmax();
  if (chain) {
    // Cluster kernels: [embed + front 0] x [back 0 + front 1] x ... x [back 7], one launch between
    // cross-attentions.  The two additive attentions only feed the mixture gates and the copy scatter,
    // so they run on a side stream: attns[0] beside the whole second stack, attns[1] beside norm1 ->
    // gen.0 -> vocabulary GEMM (fork/join by events, captured into the graph like any other edge).
    const bool fork = g_use_fork && a->h0 != nullptr && a->qa1 != nullptr && aux_ready();
    // the whole first stack (4 layers over the S0 <= 64 keys of the query memory, cross-attention included)
    // plus the first half-layer of the second stack is ONE launch; otherwise one launch per half-layer pair
    const bool stack0 = g_use_stack && a->S[0] <= case_layer_chain_max_s0();
    // attention queries, norm1 and gen.0 ride on the cluster launches as post linears (no row_linear launches)
    const bool post = g_use_post && a->Wqa_c[0] != nullptr && a->Wqa_c[1] != nullptr && a->Wg_c != nullptr;
    auto qa_post = [&](int i, float* out) {
      case_chain_post_t p;
      memset(&p, 0, sizeof(p));
      p.npost = 1; p.W = W; p.feat = a->feat;
      p.lin[0].Wc = a->Wqa_c[i]; p.lin[0].bias = a->bqa[i]; p.lin[0].out = out; p.lin[0].nchunk = 2;
      p.lin[0].seg[0] = CASE_SEG_H; p.lin[0].seg[1] = CASE_SEG_FEAT;
      return p;
    };
    int Lstart = 0;
    if (stack0) {
      float* hdst0 = fork ? a->h0 : a->h;
      void* kcs[5]; void* vcs[5]; const void* kxs[4];
      for (int l = 0; l < 5; ++l) { kcs[l] = a->kcache[l]; vcs[l] = a->vcache[l]; }
      for (int l = 0; l < 4; ++l) kxs[l] = a->Kx[l];
      case_chain_post_t p0 = qa_post(0, a->qa);
      if (xpart && g_prefetch_pct > 0 && (g_prefetch_mask & 1)) case_layer_chain_prefetch(a->Kx[4], a->xprefix, B, a->S[1], g_prefetch_pct);
      TRY(case_layer_stack(a->layers, 4, kcs, vcs, kxs, a->mask[0], W, a->S[0], nullptr, a->E, a->pe, 16.0f, a->x_in, hdst0,
                           anc, TL, a->tok, TL, a->prow, t, a->Tmax, a->bbuf, a->q2, R, 1, post ? &p0 : nullptr, st));
      if (fork) {
        CUTRY(cudaEventRecord(g_ev_fork[0], st));
        CUTRY(cudaStreamWaitEvent(g_aux, g_ev_fork[0], 0));
        TRY(stack_attention(0, hdst0, a->qa, g_aux, post));
        CUTRY(cudaEventRecord(g_ev_join[0], g_aux));
      } else {
        TRY(stack_attention(0, hdst0, a->qa, st, post));
      }
      TRY(big_xattn(4));
      Lstart = 5;
    }
    for (int L = Lstart; L <= 8; ++L) {
      const case_layer_weights_t* wb = L > 0 ? &a->layers[L - 1] : nullptr;
      const case_layer_weights_t* wf = L < 8 ? &a->layers[L] : nullptr;
      float* hdst = (L == 4 && fork) ? a->h0 : a->h;
      case_chain_post_t pl;
      memset(&pl, 0, sizeof(pl));
      if (post && L == 4) pl = qa_post(0, a->qa);
      if (post && L == 8) {
        pl = qa_post(1, fork ? a->qa1 : a->qa);
        pl.npost = 2; pl.x_in = a->x_in; pl.ln_g = a->lnN_g; pl.ln_b = a->lnN_b; pl.ln_out = a->hN;
        pl.lin[1].Wc = a->Wg_c; pl.lin[1].bias = a->bg; pl.lin[1].out = a->gfeat; pl.lin[1].nchunk = 3;
        pl.lin[1].seg[0] = CASE_SEG_XIN; pl.lin[1].seg[1] = CASE_SEG_HLN; pl.lin[1].seg[2] = CASE_SEG_FEAT;
      }
      if (xpart && g_prefetch_pct > 0 && (g_prefetch_mask & 2) && L >= 4 && L < 8) case_layer_chain_prefetch(a->Kx[L], a->xprefix, B, a->S[1], g_prefetch_pct);
      TRY(case_layer_chain(wb, wf, nullptr, a->E, a->pe, 16.0f /* sqrt(256) */, a->x_in, a->bbuf, a->part_ml,
                           a->part_acc, L > 0 ? nparts_of(L - 1) : 1, hdst, wf ? a->kcache[L] : nullptr,
                           wf ? a->vcache[L] : nullptr, anc, TL, a->tok, TL, a->prow, t, a->Tmax, a->bbuf, a->q2, R,
                           L == 0, pl.npost ? &pl : nullptr, st));
      if (L == 4 || L == 8) {
        const int i = L / 4 - 1;
        if (fork) {
          CUTRY(cudaEventRecord(g_ev_fork[i], st));
          CUTRY(cudaStreamWaitEvent(g_aux, g_ev_fork[i], 0));
          TRY(stack_attention(i, hdst, i == 0 ? a->qa : a->qa1, g_aux, post));
          if (i == 1 && g_next_prefetch && stack0) {
            PfRegions pr;
            pr.n = 0;
            auto add = [&](const void* p, size_t bytes) {
              if (p && bytes && pr.n < 16) { pr.p[pr.n] = (const char*)p; pr.bytes[pr.n] = bytes; ++pr.n; }
            };
            for (int l = 0; l < 4; ++l) {
              if (g_next_prefetch & 1) {
                add(a->layers[l].Wc, (size_t)4 * 8 * 64 * 256 * 2);
                add(a->Kx[l], (size_t)B * NH * ((a->S[0] + 63) / 64) * 8192);
              }
              if (g_next_prefetch & 2) {
                add(a->kcache[l], (size_t)R * a->Tmax * H * 2);
                add(a->vcache[l], (size_t)R * a->Tmax * H * 2);
              }
            }
            launch_k(l2_prefetch_kernel, 148, 128, 0, g_aux, pr);
            TRY(check_launch("l2_prefetch"));
          }
          CUTRY(cudaEventRecord(g_ev_join[i], g_aux));
        } else {
          TRY(stack_attention(i, hdst, a->qa, st, post));
        }
      }
      if (L == 8) break;
      TRY(big_xattn(L));
    }
    if (!post) {
      TRY(case_layernorm_rows(a->h, a->lnN_g, a->lnN_b, a->hN, R, st));
      TRY(gen0());
    }
    TRY(case_vocab_gemm(a->gfeat, a->Wv, nullptr, a->logits, R, a->V, a->ldv, dt, a->vocab_impl, a->vocab_ws, st));
    if (sparse && !getenv("CASE_SKIP_BASE")) TRY(case_vocab_base(a->logits, a->ldv, R, a->V, 0, k2, a->base_ms, a->base_e, a->base_i, st));
    if (fork) {
      CUTRY(cudaStreamWaitEvent(st, g_ev_join[0], 0));
      CUTRY(cudaStreamWaitEvent(st, g_ev_join[1], 0));
    }
  } else {
  TRY(case_embed_rows(a->E, a->pe, a->tok, TL, t, 16.0f /* sqrt(256) */, a->x_in, R, st));
  const float* hin = a->x_in;
  for (int i = 0; i < 2; ++i) {
    for (int l = 0; l < 4; ++l) {
      const int L = i * 4 + l;
      TRY(case_layer_front(hin, &a->layers[L], a->kcache[L], a->vcache[L], anc, TL, a->tok, TL, t, a->Tmax,
                           a->bbuf, a->q2, R, dt, st));
      int nparts = nparts_of(L);
      if (dt == CASE_BF16) {   // tensor-core tiles (Kx holds the interleaved K|V tiles)
        TRY(big_xattn(L));
      } else {
        TRY(case_cross_attn_partial(a->q2, a->Kx[L], a->Vx[L], a->mask[i], B, W, a->S[i], a->nsplit_x[i],
                                    a->part_ml, a->part_acc, dt, st));
      }
      TRY(case_layer_back(a->bbuf, a->part_ml, a->part_acc, nparts, &a->layers[L], a->h, R, dt, st));
      hin = a->h;
    }
    TRY(stack_attention(i, a->h, a->qa, st));
  }
  TRY(case_layernorm_rows(a->h, a->lnN_g, a->lnN_b, a->hN, R, st));
  TRY(gen0());
  TRY(case_vocab_gemm(a->gfeat, a->Wv, nullptr, a->logits, R, a->V, a->ldv, dt, a->vocab_impl, a->vocab_ws, st));
  if (sparse) TRY(case_vocab_base(a->logits, a->ldv, R, a->V, 0, k2, a->base_ms, a->base_e, a->base_i, st));
  }
  if (sparse || (g_use_tail && a->V <= case_row_tail_max_vocab())) {
    // one launch: attention merge + gates, softmax x gate, both copy scatters, top-k; the [R, V]
    // distribution is written only for the `generate` face
    case_tail_args_t ta;
    memset(&ta, 0, sizeof(ta));
    ta.R = R; ta.V = a->V; ta.W = W; ta.K = W; ta.ldl = a->ldv; ta.ldd = a->ldv; ta.mask_col0 = 0; ta.nmem = 2;
    ta.do_finalize = 1; ta.fac_ld = 2 * CASE_MAX_SPLIT; ta.map_ld = a->map_ld;
    ta.logits = a->logits; ta.hN = a->hN; ta.Wm = a->Wm; ta.bm = a->bm; ta.gates = a->gates; ta.fac = a->fac;
    ta.map = a->map;
    for (int i = 0; i < 2; ++i) {
      ta.ns[i] = a->nsplit_a[i]; ta.fac_off[i] = i * CASE_MAX_SPLIT; ta.map_off[i] = a->map_off[i]; ta.S[i] = a->S[i];
      ta.stats[i] = a->stats[i]; ta.ctxp[i] = a->ctxp[i]; ta.ctx[i] = a->ctx[i]; ta.prior[i] = a->prior[i];
      ta.attn_un[i] = a->attn_un[i];
    }
    if (a->materialize_only) { ta.dist = a->dist; } else { ta.top_vals = a->top_vals; ta.top_idx = a->top_idx; }
    ta.gate_ctx = gate ? 1 : 0;
    if (sparse && g_use_plan && a->cp_n != nullptr) {
      ta.cp_n = a->cp_n; ta.cp_uid = a->cp_uid; ta.cp_first = a->cp_first; ta.cp_start = a->cp_start; ta.cp_perm = a->cp_perm;
      ta.cp_ld = a->cp_ld;
    }
    if (sparse) {
      const bool fuse_sel = a->qcount != nullptr && g_use_fused_select;
      case_select_args_t sel = select_args(a->mode, B, W, t, a->max_len, a->Tmax, a->BOS, a->EOS, a->UNK, a->PAD,
                                           a->top_vals, a->top_idx, a->live, a->cum, a->length, a->tok, a->anc, a->parent,
                                           a->ended, a->best_key, a->best_len, a->out_tokens, a->n_live);
      TRY(case_sparse_tail(&ta, a->base_ms, a->base_e, a->base_i, k2, fuse_sel ? &sel : nullptr, a->qcount, st));
      if (fuse_sel) return 0;
    } else {
      TRY(case_row_tail(&ta, st));
    }
    if (a->materialize_only) return 0;
  } else {
    TRY(finalize());
    TRY(case_softmax_mix(a->logits, a->ldv, a->gates, a->dist, a->ldv, R, a->V, 0, st));
    for (int i = 0; i < 2; ++i) {
      TRY(case_copy_scatter(a->map, a->map_ld, a->map_off[i], a->prior[i], a->attn_un[i],
                            a->fac + (size_t)i * CASE_MAX_SPLIT, 2 * CASE_MAX_SPLIT, a->dist, a->ldv, B, W, a->S[i],
                            a->V, st));
    }
    if (a->materialize_only) return 0;
    TRY(case_topk_rows(a->dist, a->ldv, R, a->V, W, a->top_vals, a->top_idx, st));
  }
  return select_step(a->mode, B, W, t, a->max_len, a->Tmax, a->BOS, a->EOS, a->UNK, a->PAD, a->top_vals, a->top_idx,
                     a->live, a->cum, a->length, a->tok, a->anc, a->parent, a->ended, a->best_key, a->best_len,
                     a->out_tokens, a->n_live, st);
}

extern "C" int gttp_decode_step(const gttp_step_args_t* a, int t, case_stream_t stream) {
  CB_REQUIRE(a, "gttp_decode_step: null args");
  CB_REQUIRE(a->R == a->B * a->W && a->W >= 1 && a->W <= CASE_MAX_W, "gttp_decode_step: R != B*W or W out of range");
  CB_REQUIRE(t >= 0 && t < a->Tmax && a->Tmax <= CASE_MAX_T, "gttp_decode_step: t out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const int R = a->R, B = a->B, W = a->W, TL = a->Tmax + 1, dt = a->dtype;
  const float* s_in = a->state[t & 1];
  float* s_out = a->state[(t + 1) & 1];

  TRY(case_embed_rows(a->E, nullptr, a->tok, TL, t, 1.0f, a->emb, R, st));
  // two additive attentions with the previous GRU state as query (GTTP/Model.py:117-122)
  const void* Wq[2] = {a->Wqs_t, a->Wqb_t};
  const float* bq[2] = {a->bqs, a->bqb};
  const float* vv[2] = {a->vs, a->vb};
  const void* U[2] = {a->Us, a->Ub};
  const void* M[2] = {a->Ms, a->Mb};
  const uint8_t* mk[2] = {a->mask_c, a->mask_b};
  const int L[2] = {a->Lc, a->Lb};
  const int ns[2] = {a->nsplit_c, a->nsplit_b};
  for (int i = 0; i < 2; ++i) {
    case_rowlin_args_t q;
    memset(&q, 0, sizeof(q));
    q.seg[0] = seg(s_in, H, H, 1, 1);
    q.nseg = 1; q.K = H; q.Wt = Wq[i]; q.bias = bq[i]; q.N = H; q.out = a->qa; q.ldo = H;
    q.gather_idx = a->parent; q.R = R; q.dtype = dt;
    TRY(case_row_linear(&q, st));
    TRY(case_additive_attn(a->qa, U[i], M[i], vv[i], mk[i], nullptr, nullptr, 0, t, B, W, L[i], 2 * H, ns[i],
                           a->attn_un[i], a->stats[i], a->ctxp[i], a->fast_tanh, dt, st));
    TRY(case_attn_merge(a->stats[i], a->ctxp[i], ns[i], 2 * H, a->ctx[i], i == 1 ? a->fac : nullptr, CASE_MAX_SPLIT,
                        R, st));
  }
  {  // GRU cell on [emb ; src_ctx ; bg_ctx]   (Model.py:124-126)
    case_rowlin_args_t g;
    memset(&g, 0, sizeof(g));
    g.seg[0] = seg(a->emb, H, H, 1);
    g.seg[1] = seg(a->ctx[0], 2 * H, 2 * H, 1);
    g.seg[2] = seg(a->ctx[1], 2 * H, 2 * H, 1);
    g.nseg = 3; g.K = 5 * H; g.Wt = a->Wih_t; g.bias = a->bih; g.N = 3 * H; g.out = a->gi; g.ldo = 3 * H;
    g.R = R; g.dtype = dt;
    TRY(case_row_linear(&g, st));
    case_rowlin_args_t hh;
    memset(&hh, 0, sizeof(hh));
    hh.seg[0] = seg(s_in, H, H, 1, 1);
    hh.nseg = 1; hh.K = H; hh.Wt = a->Whh_t; hh.bias = a->bhh; hh.N = 3 * H; hh.out = a->gh; hh.ldo = 3 * H;
    hh.gather_idx = a->parent; hh.R = R; hh.dtype = dt;
    TRY(case_row_linear(&hh, st));
    TRY(case_gru_cell(a->gi, a->gh, s_in, a->parent, s_out, R, st));
  }
  {  // readout on [emb ; h' ; src_ctx ; bg_ctx]   (Model.py:128-130)
    case_rowlin_args_t r;
    memset(&r, 0, sizeof(r));
    r.seg[0] = seg(a->emb, H, H, 1);
    r.seg[1] = seg(s_out, H, H, 1);
    r.seg[2] = seg(a->ctx[0], 2 * H, 2 * H, 1);
    r.seg[3] = seg(a->ctx[1], 2 * H, 2 * H, 1);
    r.nseg = 4; r.K = 6 * H; r.Wt = a->Wr_t; r.bias = a->br; r.N = H; r.out = a->feat; r.ldo = H;
    r.R = R; r.dtype = dt;
    TRY(case_row_linear(&r, st));
  }
  TRY(case_vocab_gemm(a->feat, a->Wv, a->bv, a->logits, R, a->V, a->ldv, dt, a->vocab_impl, a->vocab_ws, st));
  TRY(case_gttp_gates(a->feat, a->wc, a->bc, a->gates, a->fac, CASE_MAX_SPLIT, ns[1], R, st));
  if (g_use_tail && a->V <= case_row_tail_max_vocab()) {
    case_tail_args_t ta;
    memset(&ta, 0, sizeof(ta));
    ta.R = R; ta.V = a->V; ta.W = W; ta.K = W; ta.ldl = a->ldv; ta.ldd = a->ldv; ta.mask_col0 = 1; ta.nmem = 1;
    ta.do_finalize = 0; ta.fac_ld = CASE_MAX_SPLIT; ta.map_ld = a->map_ld;
    ta.logits = a->logits; ta.gates = a->gates; ta.fac = a->fac; ta.map = a->map;
    ta.S[0] = a->Lb; ta.attn_un[0] = a->attn_un[1];
    if (a->materialize_only) { ta.dist = a->dist; } else { ta.top_vals = a->top_vals; ta.top_idx = a->top_idx; }
    TRY(case_row_tail(&ta, st));
    if (a->materialize_only) return 0;
  } else {
    TRY(case_softmax_mix(a->logits, a->ldv, a->gates, a->dist, a->ldv, R, a->V, 1, st));
    TRY(case_copy_scatter(a->map, a->map_ld, 0, nullptr, a->attn_un[1], a->fac, CASE_MAX_SPLIT, a->dist, a->ldv, B, W,
                          a->Lb, a->V, st));
    if (a->materialize_only) return 0;
    TRY(case_topk_rows(a->dist, a->ldv, R, a->V, W, a->top_vals, a->top_idx, st));
  }
  return select_step(a->mode, B, W, t, a->max_len, a->Tmax, a->BOS, a->EOS, a->UNK, a->PAD, a->top_vals, a->top_idx,
                     a->live, a->cum, a->length, a->tok, a->anc, a->parent, a->ended, a->best_key, a->best_len,
                     a->out_tokens, a->n_live, st);
}

/* layout self-check for FFI bindings: sizeof of each argument struct, by index */
extern "C" size_t case_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(case_seg_t);
    case 1: return sizeof(case_rowlin_args_t);
    case 2: return sizeof(case_layer_weights_t);
    case 3: return sizeof(case_select_args_t);
    case 4: return sizeof(case_step_args_t);
    case 5: return sizeof(gttp_step_args_t);
    case 6: return sizeof(case_tail_args_t);
    case 7: return sizeof(case_chain_post_t);
    default: return 0;
  }
}
