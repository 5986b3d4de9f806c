"""The torch custom-op layer (TORCH_LIBRARY(case_b200, ...), case_rg_b200/csrc/torch_ops.cpp) over the C ABI: each op
against plain torch on the same inputs, dispatcher-level argument checking, and the whole decode driven through
case_b200::decode_step instead of ctypes (CUDA-graph capture included)."""
import pytest
import torch

from case_rg_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
H = 256


@pytest.fixture(scope='module')
def ops():
    from case_rg_b200 import _lib as L
    return L.load_torch_ops()


def test_topk_rows_op_matches_torch_and_rejects_cpu_tensors(ops):
    g = torch.Generator().manual_seed(3)
    x = torch.rand(19, 30528, generator=g)
    x[:, ::9] = 0.25
    vals, idx = ops.topk_rows(x.to(DEV), 30522, 4)
    key = x[:, :30522].double() - torch.arange(30522).double() * 1e-12          # value desc, index asc
    want = key.topk(4, dim=1).indices
    assert idx.dtype == torch.int32 and torch.equal(idx.cpu().long(), want)
    assert torch.equal(vals.cpu(), x[:, :30522].gather(1, want))
    with pytest.raises(NotImplementedError):          # no CPU kernel is registered: there is no CPU fallback
        ops.topk_rows(x, 30522, 4)
    with pytest.raises(RuntimeError):                 # dtype checked at the op boundary
        ops.topk_rows(x.to(DEV).double(), 30522, 4)
    with pytest.raises(RuntimeError):
        ops.topk_rows(x.to(DEV), 30522, 9)


def test_softmax_mix_and_copy_scatter_ops_equal_the_reference_formula(ops):
    """dist = g0 * softmax(logits) + onehot^T . (F * prior * exp(e - M))   (CaSE/Model.py:41-43)."""
    B, W, S, V = 3, 2, 200, 777
    ld = 784
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(B * W, ld, generator=g) * 2).to(DEV)
    gates = torch.rand(B * W, 4, generator=g).to(DEV)
    mp = torch.randint(0, V, (B, S), generator=g, dtype=torch.int32).to(DEV)
    e = torch.randn(B * W, S, generator=g).to(DEV)
    e[:, ::7] = float('-inf')
    prior = torch.rand(B, S, generator=g).to(DEV)
    fac = torch.zeros(B * W, 16, device=DEV)
    fac[:, 0] = torch.rand(B * W, generator=g).to(DEV) * 0.1
    fac[:, 1] = e.max(1).values
    dist = ops.softmax_mix(logits, gates, V, False)
    assert torch.allclose(dist[:, :V], gates[:, :1] * torch.softmax(logits[:, :V], 1), rtol=1e-5, atol=1e-8)
    out = ops.copy_scatter_(dist, mp, 0, prior, e, fac, W, V)
    assert out.data_ptr() == dist.data_ptr()          # in place (Tensor(a!))
    coef = fac[:, :1] * prior.repeat_interleave(W, 0) * torch.exp(e - fac[:, 1:2])
    oh = torch.zeros(B, S, V, device=DEV).scatter_(2, mp.long().unsqueeze(2), 1.0)
    want = gates[:, :1] * torch.softmax(logits[:, :V], 1) + torch.bmm(coef.view(B, W, S), oh).view(B * W, V)
    assert float((dist[:, :V] - want).abs().max() / want.abs().max()) < 1e-5
    with pytest.raises(RuntimeError):
        ops.copy_scatter_(dist, mp.long(), 0, prior, e, fac, W, V)     # map must be int32


def test_vocab_gemm_op_tcgen05_and_simt(ops):
    from case_rg_b200.engine import pack_vocab_tc
    V, R = 30522, 256
    g = torch.Generator().manual_seed(7)
    f = torch.randn(R, H, generator=g).to(DEV)
    Wv = (torch.randn(V, H, generator=g) * 0.05).to(DEV)
    want = f.bfloat16().float() @ Wv.bfloat16().float().t()
    got_tc = ops.vocab_gemm(f, pack_vocab_tc(Wv), None, V, 1)
    assert float((got_tc[:, :V] - want).abs().max() / want.abs().max()) < 2e-3
    got = ops.vocab_gemm(f, Wv.contiguous(), None, V, 0)
    assert float((got[:, :V] - f @ Wv.t()).abs().max() / want.abs().max()) < 1e-5


@pytest.mark.parametrize('family', ['case', 'gttp'])
def test_decode_through_the_custom_op_equals_the_ctypes_path(ops, family):
    """use_torch_ops = True: every step (eager and captured into the CUDA graph) goes through case_b200::decode_step /
    gttp_step; the answers are bit-identical to the ctypes path (same launcher, same argument block)."""
    from case_rg_b200 import generations as FG
    T, W = 10, 4
    if family == 'case':
        V, B = 3000, 6
        sd = syn.make_case_decoder_state(61, V, H, peaked=0.3, boost={syn.EOS: 8.0}, gen_gate_bias=2.0)
        d = syn.make_case_inputs(62, B, 24, 3, 50, V, H).to(DEV)
        data = dict(mem_q=d.mem_q, mem_p=d.mem_p, query=d.query, passage=d.passage, prior_q=d.prior_q,
                    prior_p=d.prior_p, answer_rep=d.answer_rep, source_map=d.source_map)
        mk = lambda graph: FG.FastCaSE(sd, device=DEV, dtype='bf16', use_graph=graph)
    else:
        V, B = 4000, 5
        sd = syn.make_gttp_state(71, V, H, H, peaked=0.3, boost={syn.EOS: 4.0})
        d = syn.make_gttp_inputs(72, B, 20, 3, 30, V, H).to(DEV)
        data = dict(context=d.context, background=d.background, background_map=d.background_map,
                    src_output=d.src_output, bg_output=d.bg_output, init_state=d.init_state)
        mk = lambda graph: FG.FastGTTP(sd, device=DEV, dtype='fp32', use_graph=graph)
    want = FG.beam(mk(True), data, None, T, W).cpu()
    for graph in (False, True):
        model = mk(graph)
        FG.beam(model, data, None, 2, W)                       # build an engine, then switch its step calls to the op
        type(model.last_engine).use_torch_ops = True
        try:
            got = FG.beam(model, data, None, T, W).cpu()
        finally:
            type(model.last_engine).use_torch_ops = False
        assert torch.equal(got, want), (graph, got, want)
