// Torch custom-op layer over the C ABI of libcase_b200.so: TORCH_LIBRARY(case_b200, ...).
//
// Thin by construction: every op validates device / dtype / shape / contiguity / alignment at the
// dispatcher boundary, allocates its outputs with the caching allocator, takes the stream from
// at::cuda::getCurrentCUDAStream() and calls ONE launcher of include/case_b200.h; a non-zero return code becomes a C++
// exception carrying case_last_error().  Only CUDA kernels are registered: calling an op with CPU tensors fails in the
// dispatcher ("no kernel for backend CPU") - there is no CPU fallback.  `decode_step` / `gttp_step` take the argument
// block of the step orchestrators (a byte blob owned by the engine: pointers to buffers the engine keeps alive) and are
// registered for every backend, since the blob itself lives in host memory.
//
// Reference call sites the ops stand for: Utils.topk (common/Utils.py:156-168), the one-hot copy bmm (CaSE/Model.py:43,
// GTTP/Model.py:37-40), gen softmax x gate (Model.py:34,41), gen.2 / gen.linear (Model.py:34, GTTP/Model.py:8),
// nn.MultiheadAttention over the passage memory (TransformerDecoder.py:81), BilinearAttention (BilinearAttention.py:24-60),
// the eval loop body (Model.py:94-122) and the GTTP step (GTTP/Model.py:176-193).
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>

#include "../../include/case_b200.h"

namespace {

using at::Tensor;

inline void rc_check(int rc, const char* what) {
  TORCH_CHECK(rc == 0, "case_b200::", what, " failed (code ", rc, "): ", case_last_error());
}
inline void need(const Tensor& t, const char* name, at::ScalarType dt, int64_t dim = -1) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (there is no CPU fallback)");
  TORCH_CHECK(t.scalar_type() == dt, name, " has dtype ", t.scalar_type(), ", expected ", dt);
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
  TORCH_CHECK(dim < 0 || t.dim() == dim, name, " must have ", dim, " dimensions, got ", t.dim());
}
inline case_stream_t cur_stream(const Tensor& t) {
  return (case_stream_t)c10::cuda::getCurrentCUDAStream(t.get_device()).stream();
}
inline const void* optp(const c10::optional<Tensor>& t) { return t.has_value() ? t->data_ptr() : nullptr; }

// values desc, ties -> lower index first (Utils.topk)
std::tuple<Tensor, Tensor> topk_rows(const Tensor& dist, int64_t V, int64_t k) {
  need(dist, "dist", at::kFloat, 2);
  TORCH_CHECK(V >= 1 && V <= dist.size(1) && k >= 1 && k <= CASE_MAX_W, "topk_rows: need 1 <= V <= ld and 1 <= k <= 8");
  c10::cuda::CUDAGuard g(dist.device());
  Tensor vals = at::empty({dist.size(0), k}, dist.options());
  Tensor idx = at::empty({dist.size(0), k}, dist.options().dtype(at::kInt));
  rc_check(case_topk_rows(dist.data_ptr<float>(), (int)dist.size(1), (int)dist.size(0), (int)V, (int)k, vals.data_ptr<float>(),
                          idx.data_ptr<int32_t>(), cur_stream(dist)), "topk_rows");
  return {vals, idx};
}

// dist[r, map[b, off + s]] += F[r] * prior[b, s] * exp(e[r, s] - M[r]),  (F, M) = fac[r, 0..1]
Tensor copy_scatter_(Tensor dist, const Tensor& map, int64_t map_off, const c10::optional<Tensor>& prior, const Tensor& attn_un,
                     const Tensor& fac, int64_t W, int64_t V) {
  need(dist, "dist", at::kFloat, 2); need(map, "map", at::kInt, 2); need(attn_un, "attn_un", at::kFloat, 2);
  need(fac, "fac", at::kFloat, 2);
  if (prior.has_value()) need(*prior, "prior", at::kFloat, 2);
  const int64_t R = dist.size(0), B = map.size(0), S = attn_un.size(1);
  TORCH_CHECK(W >= 1 && R == B * W && attn_un.size(0) == R && fac.size(0) == R && fac.size(1) >= 2, "copy_scatter_: R != B*W or fac too narrow");
  TORCH_CHECK(map_off >= 0 && map_off + S <= map.size(1) && V <= dist.size(1), "copy_scatter_: map / dist too small");
  TORCH_CHECK(!prior.has_value() || (prior->size(0) == B && prior->size(1) == S), "copy_scatter_: prior must be [B, S]");
  c10::cuda::CUDAGuard g(dist.device());
  rc_check(case_copy_scatter(map.data_ptr<int32_t>(), (int)map.size(1), (int)map_off, (const float*)optp(prior),
                             attn_un.data_ptr<float>(), fac.data_ptr<float>(), (int)fac.size(1), dist.data_ptr<float>(),
                             (int)dist.size(1), (int)B, (int)W, (int)S, (int)V, cur_stream(dist)), "copy_scatter_");
  return dist;
}

// gates[r][0] * softmax(logits[r, :V])
Tensor softmax_mix(const Tensor& logits, const Tensor& gates, int64_t V, bool mask_col0) {
  need(logits, "logits", at::kFloat, 2); need(gates, "gates", at::kFloat, 2);
  TORCH_CHECK(gates.size(0) == logits.size(0) && gates.size(1) == 4 && V >= 1 && V <= logits.size(1), "softmax_mix: gates must be [R, 4], V <= ld");
  c10::cuda::CUDAGuard g(logits.device());
  Tensor dist = at::zeros_like(logits);
  rc_check(case_softmax_mix(logits.data_ptr<float>(), (int)logits.size(1), gates.data_ptr<float>(), dist.data_ptr<float>(),
                            (int)dist.size(1), (int)logits.size(0), (int)V, mask_col0 ? 1 : 0, cur_stream(logits)), "softmax_mix");
  return dist;
}

// logits[R, ld] = f[R, 256] . Wv^T (+ bias); impl 1 = tcgen05 kernel on the packed bf16 weight
Tensor vocab_gemm(const Tensor& f, const Tensor& Wv, const c10::optional<Tensor>& bias, int64_t V, int64_t impl) {
  need(f, "f", at::kFloat, 2);
  TORCH_CHECK(f.size(1) == CASE_H && Wv.is_cuda() && Wv.is_contiguous(), "vocab_gemm: f must be [R, 256], Wv a contiguous CUDA tensor");
  const bool bf = Wv.scalar_type() == at::kBFloat16;
  TORCH_CHECK(bf || Wv.scalar_type() == at::kFloat, "vocab_gemm: Wv must be bf16 or fp32");
  TORCH_CHECK(impl == 0 || (impl == 1 && bf && (size_t)Wv.numel() * 2 >= case_vocab_tc_packed_weight_bytes((int)V)),
              "vocab_gemm: impl 1 needs the packed bf16 weight (case_vocab_tc_packed_weight_bytes)");
  if (bias.has_value()) need(*bias, "bias", at::kFloat, 1);
  c10::cuda::CUDAGuard g(f.device());
  const int64_t R = f.size(0), ld = (V + 7) / 8 * 8;
  Tensor logits = at::empty({R, ld}, f.options());
  Tensor ws = at::empty({(int64_t)std::max<size_t>(16, case_vocab_tc_workspace_bytes((int)R))}, f.options().dtype(at::kByte));
  rc_check(case_vocab_gemm(f.data_ptr<float>(), Wv.data_ptr(), (const float*)optp(bias), logits.data_ptr<float>(), (int)R, (int)V,
                           (int)ld, bf ? CASE_BF16 : CASE_F32, (int)impl, ws.data_ptr(), cur_stream(f)), "vocab_gemm");
  return logits;
}

// passage cross-attention over the compacted K|V tile stream -> flash-decoding partials
std::tuple<Tensor, Tensor> cross_attn_part(const Tensor& q2, const Tensor& KV, const Tensor& ncount, const Tensor& tile_prefix,
                                           int64_t W, int64_t S) {
  need(q2, "q2", at::kFloat, 2); need(KV, "KV", at::kBFloat16); need(ncount, "ncount", at::kInt, 1);
  need(tile_prefix, "tile_prefix", at::kInt, 1);
  const int64_t B = ncount.size(0), R = q2.size(0);
  TORCH_CHECK(q2.size(1) == CASE_H && W >= 1 && R == B * W && tile_prefix.size(0) == B + 1, "cross_attn_part: q2 [B*W, 256], tile_prefix [B+1]");
  TORCH_CHECK(KV.numel() >= B * CASE_NH * ((S + 63) / 64) * 2 * 64 * CASE_HD, "cross_attn_part: KV smaller than [B][NH][ceil(S/64)][2][64][32]");
  c10::cuda::CUDAGuard g(q2.device());
  const int nslot = case_cross_attn_part_slots((int)S);
  Tensor ml = at::empty({R, CASE_NH, nslot, 2}, q2.options());
  Tensor acc = at::empty({R, CASE_NH, nslot, CASE_HD}, q2.options());
  rc_check(case_cross_attn_part(q2.data_ptr<float>(), KV.data_ptr(), ncount.data_ptr<int32_t>(), tile_prefix.data_ptr<int32_t>(),
                                (int)B, (int)W, (int)S, nslot, ml.data_ptr<float>(), acc.data_ptr<float>(), cur_stream(q2)),
           "cross_attn_part");
  return {ml, acc};
}

// gate-form additive attention: raw masked scores, per-split softmax statistics, per-split gate partials
std::tuple<Tensor, Tensor, Tensor> additive_attn_gate(const Tensor& qa, const Tensor& U, const Tensor& G, const Tensor& v,
                                                      const Tensor& mask, const c10::optional<Tensor>& prior, int64_t W,
                                                      int64_t nsplit, bool fast_tanh) {
  need(qa, "qa", at::kFloat, 2); need(U, "U", at::kBFloat16, 3); need(G, "G", at::kFloat, 3); need(v, "v", at::kFloat, 1);
  need(mask, "mask", at::kByte, 2);
  if (prior.has_value()) need(*prior, "prior", at::kFloat, 2);
  const int64_t B = U.size(0), S = U.size(1), R = qa.size(0);
  TORCH_CHECK(U.size(2) == CASE_H && qa.size(1) == CASE_H && v.size(0) == CASE_H && R == B * W, "additive_attn_gate: H = 256, R = B*W");
  TORCH_CHECK(G.size(0) == B && G.size(1) == S && G.size(2) == 4 && mask.size(0) == B && mask.size(1) == S, "additive_attn_gate: G [B,S,4], mask [B,S]");
  TORCH_CHECK(nsplit >= 1 && nsplit <= CASE_MAX_SPLIT, "additive_attn_gate: 1 <= nsplit <= 16");
  c10::cuda::CUDAGuard g(qa.device());
  Tensor e = at::empty({R, S}, qa.options());
  Tensor stats = at::empty({R, nsplit, 4}, qa.options());
  Tensor gp = at::empty({R, nsplit, 4}, qa.options());
  rc_check(case_additive_attn_gate(qa.data_ptr<float>(), U.data_ptr(), G.data_ptr<float>(), v.data_ptr<float>(),
                                   mask.data_ptr<uint8_t>(), (const float*)optp(prior), nullptr, 0, 0, (int)B, (int)W, (int)S,
                                   (int)nsplit, e.data_ptr<float>(), stats.data_ptr<float>(), gp.data_ptr<float>(),
                                   fast_tanh ? 1 : 0, nullptr, nullptr, nullptr, nullptr, cur_stream(qa)), "additive_attn_gate");
  return {e, stats, gp};
}

// whole-step orchestrators: `args` = the engine's argument block as a byte tensor in HOST memory
void decode_step(const Tensor& args, int64_t t) {
  TORCH_CHECK(args.device().is_cpu() && args.scalar_type() == at::kByte && args.is_contiguous() &&
              (size_t)args.numel() == case_struct_size(4), "decode_step: args must be the case_step_args_t blob (uint8, host)");
  rc_check(case_decode_step((const case_step_args_t*)args.data_ptr(), (int)t, (case_stream_t)c10::cuda::getCurrentCUDAStream().stream()),
           "decode_step");
}
void gttp_step(const Tensor& args, int64_t t) {
  TORCH_CHECK(args.device().is_cpu() && args.scalar_type() == at::kByte && args.is_contiguous() &&
              (size_t)args.numel() == case_struct_size(5), "gttp_step: args must be the gttp_step_args_t blob (uint8, host)");
  rc_check(gttp_decode_step((const gttp_step_args_t*)args.data_ptr(), (int)t, (case_stream_t)c10::cuda::getCurrentCUDAStream().stream()),
           "gttp_step");
}

}  // namespace

TORCH_LIBRARY(case_b200, m) {
  m.def("topk_rows(Tensor dist, int V, int k) -> (Tensor, Tensor)");
  m.def("copy_scatter_(Tensor(a!) dist, Tensor map, int map_off, Tensor? prior, Tensor attn_un, Tensor fac, int W, int V) -> Tensor(a!)");
  m.def("softmax_mix(Tensor logits, Tensor gates, int V, bool mask_col0) -> Tensor");
  m.def("vocab_gemm(Tensor f, Tensor Wv, Tensor? bias, int V, int impl) -> Tensor");
  m.def("cross_attn_part(Tensor q2, Tensor KV, Tensor ncount, Tensor tile_prefix, int W, int S) -> (Tensor, Tensor)");
  m.def("additive_attn_gate(Tensor qa, Tensor U, Tensor G, Tensor v, Tensor mask, Tensor? prior, int W, int nsplit, bool fast_tanh) -> (Tensor, Tensor, Tensor)");
  m.def("decode_step(Tensor args, int t) -> ()");
  m.def("gttp_step(Tensor args, int t) -> ()");
}

TORCH_LIBRARY_IMPL(case_b200, CUDA, m) {
  m.impl("topk_rows", &topk_rows);
  m.impl("copy_scatter_", &copy_scatter_);
  m.impl("softmax_mix", &softmax_mix);
  m.impl("vocab_gemm", &vocab_gemm);
  m.impl("cross_attn_part", &cross_attn_part);
  m.impl("additive_attn_gate", &additive_attn_gate);
}

TORCH_LIBRARY_IMPL(case_b200, CompositeExplicitAutograd, m) {
  m.impl("decode_step", &decode_step);
  m.impl("gttp_step", &gttp_step);
}
