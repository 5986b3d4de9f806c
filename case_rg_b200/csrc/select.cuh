// Search bookkeeping on the device: greedy token hand-off and the beam step of
// common/Generations.py:136-185, with no host round trips (shared by the standalone kernel of select.cu
// and the fused form at the end of sparse_tail.cu).
#pragma once
#include "common.cuh"

namespace cb {

// One warp runs the bookkeeping of query b (all 32 lanes must call).  top_vals / top_idx may have been
// written by other CTAs of the same launch (fused form: the last CTA of the query's W rows calls this
// after a fence + atomic), hence the L2 (.cg) loads.
__device__ __forceinline__ void beam_select_query(const case_select_args_t& a, int b, int lane) {
  const int W = a.W, t = a.t, TL = a.Tmax + 1;
  const int r0 = b * W;
  // extended vocabulary: ids >= V_in are dynamic (OOV) entries - the decoder is fed UNK, tok_ext keeps the id
  const bool ext = a.V_in > 0 && a.tok_ext != nullptr;
  int32_t* tok_o = ext ? a.tok_ext : a.tok;
  auto feed = [&](int id) { return (ext && id >= a.V_in) ? a.UNK : id; };

  if (a.mode != CASE_MODE_BEAM) {   // W == 1
    if (lane == 0) {
      int tk = __ldcg(a.top_idx + r0);
      if (a.mode == CASE_MODE_PROTO_GREEDY) {
        const int was_ended = a.ended[b];
        const int this_end = tk == a.EOS;
        if (t == 0) { if (this_end) tk = a.UNK; }          // Generations.py:99-100
        else if (was_ended) tk = a.PAD;                    // Generations.py:101-102
        a.ended[b] = was_ended | this_end;
      }
      a.out_tokens[(size_t)b * a.Tmax + t] = tk;
      a.tok[(size_t)r0 * TL + t + 1] = feed(tk);
      if (ext) tok_o[(size_t)r0 * TL + t + 1] = tk;
      a.parent[r0] = r0;
      a.live[r0] = 1;
      a.n_live[b] = 1;
    }
    for (int j = lane; j <= t; j += 32) a.anc_out[(size_t)r0 * TL + j] = r0;
    return;
  }

  __shared__ double key[64], ccum[64];
  __shared__ int ctok[64], cpar[64], clen[64], order[CASE_MAX_W], dst[CASE_MAX_W];
  __shared__ int s_nlive, s_best;
  int nl = 0;
  for (int w = 0; w < W; ++w) nl += a.live[r0 + w] != 0;
  const int ncand = nl * W;
  for (int c = lane; c < ncand; c += 32) {
    const int w = c / W, j = c % W, r = r0 + w;
    const double p = (double)__ldcg(a.top_vals + (size_t)r * W + j);
    const double cm = a.cum[r] + (-log(p + 1e-10));         // Generations.py:170, Node.cum_cost :198
    const int ln = a.length[r] + 1;                         // Node.length :199
    key[c] = cm / (double)ln;
    ccum[c] = cm;
    ctok[c] = __ldcg(a.top_idx + (size_t)r * W + j);
    cpar[c] = r;
    clen[c] = ln;
  }
  __syncwarp();
  // stable ascending rank (sorted(..., key=cum_cost/length)[:width], Generations.py:180)
  for (int c = lane; c < ncand; c += 32) {
    const double k = key[c];
    int rank = 0;
    for (int o = 0; o < ncand; ++o) rank += (key[o] < k) || (key[o] == k && o < c);
    if (rank < W) order[rank] = c;
  }
  __syncwarp();
  const int nsel = min(W, ncand);
  if (lane == 0) {
    int nn = 0, bc = -1;
    double bk = a.best_key[b];
    for (int k = 0; k < nsel; ++k) {
      const int c = order[k];
      const bool fin = (ctok[c] == a.EOS) || (t == a.max_len - 1);   // Generations.py:139
      if (fin) {
        if (key[c] < bk) { bk = key[c]; bc = c; }                     // first finisher wins ties (:184)
        dst[k] = -1;
      } else {
        dst[k] = nn++;
      }
    }
    s_nlive = nn;
    s_best = bc;
    if (bc >= 0) a.best_key[b] = bk;
    a.n_live[b] = nn;
  }
  __syncwarp();
  // read everything the new slots need from the old state before any slot is overwritten
  // (anc is double-buffered; cum/length/live were staged in shared memory above)
  for (int k = 0; k < nsel; ++k) {
    if (dst[k] < 0) continue;
    const int c = order[k], rn = r0 + dst[k], pr = cpar[c];
    for (int j = lane; j < t; j += 32) a.anc_out[(size_t)rn * TL + j] = a.anc_in[(size_t)pr * TL + j];
    if (lane == 0) {
      a.anc_out[(size_t)rn * TL + t] = pr;
      a.tok[(size_t)rn * TL + t + 1] = feed(ctok[c]);
      if (ext) tok_o[(size_t)rn * TL + t + 1] = ctok[c];
      a.cum[rn] = ccum[c];
      a.length[rn] = clen[c];
      a.live[rn] = 1;
      a.parent[rn] = pr;
    }
  }
  for (int w = s_nlive + lane; w < W; w += 32) {     // dead slots: harmless, in-range contents
    const int rn = r0 + w;
    a.live[rn] = 0;
    a.parent[rn] = rn;
    a.tok[(size_t)rn * TL + t + 1] = a.PAD;
    if (ext) tok_o[(size_t)rn * TL + t + 1] = a.PAD;
    for (int j = 0; j <= t; ++j) a.anc_out[(size_t)rn * TL + j] = rn;
  }
  if (s_best >= 0) {   // record the new best finished sequence: BOS dropped, EOS kept (:188)
    const int c = s_best, pr = cpar[c];
    int* out = a.out_tokens + (size_t)b * a.Tmax;
    for (int j = lane; j < a.max_len; j += 32) {
      int v = a.PAD;
      if (j < t) {
        const int src = (j + 1 == t) ? pr : a.anc_in[(size_t)pr * TL + j + 1];
        v = tok_o[(size_t)src * TL + j + 1];
      } else if (j == t) {
        v = ctok[c];
      }
      out[j] = v;
    }
    if (lane == 0) a.best_len[b] = t + 1;
  }
}

}  // namespace cb
