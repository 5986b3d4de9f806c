"""SURVEY.md §8f N4 (GLKS Mixturer on the a6 / a7 / a9 kernels) and the extended (vocab + OOV) distribution:
``FastMixturer`` / ``copy_topk`` against vectors from the unmodified reference (tests/golden/make_glks_golden.py), the
GLKS face installed into a LIVE reference GLKS model, and the CaSE search with a per-query dynamic vocabulary behind the
fixed one (``n_oov``) against the oracle's extended-vocabulary stepper."""
import copy
import os

import numpy as np
import pytest
import torch

from case_rg_b200 import synthetic as syn
from helpers import GOLDEN, H, glks_inputs

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_fast_mixturer_and_topk_match_reference_golden():
    from case_rg_b200.glks import FastMixturer, topk_rows
    x = glks_inputs()
    z = np.load(os.path.join(GOLDEN, 'glks_mixturer.npz'))
    V = x['p_v'].size(1)
    mix = FastMixturer(H)
    mix.load_state_dict({'linear1.weight': x['w'], 'linear1.bias': x['b']}, strict=True)
    mix = mix.to(DEV)
    d = {k: v.to(DEV) for k, v in x.items()}
    with torch.no_grad():
        p = mix(d['state'], d['p_v'], d['p_k'], d['bmap'])                                    # index form
        oh = torch.zeros(d['bmap'].size(0), d['bmap'].size(1), V, device=DEV).scatter_(2, d['bmap'].unsqueeze(2), 1.0)
        p2 = mix(d['state'], d['p_v'], d['p_k'], oh)                                          # the reference's one-hot form
    assert torch.equal(p, p2)
    np.testing.assert_allclose(p.cpu().numpy(), z['p'], rtol=2e-5, atol=1e-9)
    for k, name in ((4, 'top4'), (1, 'top1')):
        v, i = topk_rows(p, k)
        assert i.dtype == torch.int64 and np.array_equal(i.cpu().numpy(), z[name + '_i'])
        np.testing.assert_allclose(v.cpu().numpy(), z[name + '_v'], rtol=2e-5)
    with pytest.raises(RuntimeError):
        mix.cpu()(x['state'], x['p_v'], x['p_k'], x['bmap'])          # no CPU fallback


def test_copy_topk_matches_reference_golden():
    """Utils.copy_topk (Utils.py:170-178): dynamic entries folded onto their vocabulary ids (two OOV words onto UNK),
    in-vocabulary dynamic entries zeroed, top-k over V + D - indices beyond V are selected."""
    from case_rg_b200.glks import copy_topk
    x = glks_inputs()
    z = np.load(os.path.join(GOLDEN, 'glks_mixturer.npz'))
    V = x['p_v'].size(1)
    d = {k: v.to(DEV) for k, v in x.items()}
    v, i = copy_topk(d['gen_ext'], d['vmap'], d['overlap'], 5)
    assert np.array_equal(i.cpu().numpy(), z['copy5_i']) and int(i.max()) >= V
    np.testing.assert_allclose(v.cpu().numpy(), z['copy5_v'], rtol=1e-5)
    oh = torch.zeros(d['vmap'].size(0), d['vmap'].size(1), V, device=DEV).scatter_(2, d['vmap'].unsqueeze(2), 1.0)
    v2, i2 = copy_topk(d['gen_ext'], oh, d['overlap'], 5)             # one-hot form of the reference
    assert torch.equal(i, i2)


def test_install_fast_glks_into_live_reference_glks():
    """GLKS.forward(data, 'test') (Generations.greedy over the unmodified model, GLKS/Model.py:255-262) before and after
    install_fast_glks: same answers.  The reference model stays on the CPU (its GRU encoders pass device-side lengths to
    pack_padded_sequence); generator + softmax, Mixturer and top-k run on cuda:0."""
    from baseline import refshim
    from case_rg_b200.glks import install_fast_glks, FastMixturer
    if refshim.reference_root() is None:
        pytest.skip('reference snapshot baseline/_ref not present')
    ns = refshim.load_reference()
    V, T, B = 1500, 8, 3
    vocab2id, id2vocab = syn.make_vocab(V)
    torch.manual_seed(21)
    model = ns.glks.GLKS(4, 2, H, H, vocab2id, id2vocab, T, 1).eval()
    with torch.no_grad():
        model.v_generator.generator.weight.mul_(6.0)                   # margins well above fp32 summation-order noise
    ginp = syn.make_gttp_inputs(22, B, 12, 3, 16, V, H)
    data = {'id': ginp.ids, 'context': ginp.context, 'background': ginp.background, 'background_map': ginp.background_map}
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: False      # reference helpers jump to CUDA whenever it is visible (Utils.py:300-303)
    try:
        with torch.no_grad():
            want = model(copy.copy(data), method='test')['answer']
            install_fast_glks(model, device=DEV)
            assert isinstance(model.mixture, FastMixturer)
            got = model(copy.copy(data), method='test')['answer']
    finally:
        torch.cuda.is_available = real
    assert got.shape == want.shape == (B, T) and torch.equal(got, want), (got, want)
    assert set(model.state_dict().keys()) == set(ns.glks.GLKS(4, 2, H, H, vocab2id, id2vocab, T, 1).state_dict().keys())


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_case_search_over_extended_vocabulary_vs_oracle(dtype):
    """n_oov > 0: a fifth of the source positions point at per-query dynamic words (ids V .. V + n_oov - 1).  The sparse
    tail ranks them by their copy mass alone, select feeds them back as UNK and returns the extended ids; greedy and
    beam answers equal the oracle's extended-vocabulary stepper (fp32 storage: identical; bf16: first tokens), the
    materialised distribution has V + n_oov columns, sums to 1 and matches the oracle's."""
    from case_rg_b200 import generations as FG
    from oracle import generations as OG
    from oracle.case_decoder import CaseOracle
    V, B, T, W, NO = 3000, 6, 8, 4, 50
    sd = syn.make_case_decoder_state(41, V, H, peaked=0.3, boost={syn.EOS: 6.0}, gen_gate_bias=0.5)
    inp = syn.make_case_inputs(42, B, 24, 4, 40, V, H)
    g = torch.Generator().manual_seed(43)
    oov = torch.rand(inp.source_map.shape, generator=g) < 0.2
    oov &= inp.source_map > 102                                        # content words only, specials stay in the vocabulary
    inp.source_map = torch.where(oov, V + torch.randint(0, NO, inp.source_map.shape, generator=g), inp.source_map)
    d = inp.to(DEV)
    data = dict(mem_q=d.mem_q, mem_p=d.mem_p, query=d.query, passage=d.passage, prior_q=d.prior_q, prior_p=d.prior_p,
                answer_rep=d.answer_rep, source_map=d.source_map)
    orc = CaseOracle(sd)
    model = FG.FastCaSE(sd, device=DEV, dtype=dtype, n_oov=NO)
    want_g = OG.greedy(orc.incremental(inp, n_oov=NO), T)
    want_b = OG.beam(orc.incremental(inp, n_oov=NO), T, W)
    assert int(want_b.max()) >= V or int(want_g.max()) >= V, 'the case must actually select dynamic words'
    got_g = FG.greedy(model, data, None, T).cpu()
    got_b = FG.beam(model, data, None, T, W).cpu()
    if dtype == 'fp32':
        assert torch.equal(got_g, want_g), (got_g, want_g)
        assert torch.equal(got_b, want_b), (got_b, want_b)
    else:
        assert float((got_g[:, :2] == want_g[:, :2]).float().mean()) >= 0.8
        assert int(got_b.max()) < V + NO
    # the `generate` face: [R, V + n_oov] distribution of step 0
    eng = model.last_engine
    eng.state.reset()
    live = eng.state.live.bool().clone()
    dist = eng.step_distribution(0)
    torch.cuda.synchronize()
    assert dist.shape[1] == V + NO and torch.allclose(dist.sum(1)[live], torch.ones(B, device=DEV), atol=2e-3)
    st = orc.incremental(inp, n_oov=NO)
    ref = st.advance(torch.arange(B), torch.full((B,), syn.BOS))
    err = float((dist[live].cpu() - ref).abs().max() / ref.abs().max())
    assert err < (1e-4 if dtype == 'fp32' else 8e-2), err
    assert float(ref[:, V:].sum()) > 0 and float(dist[live][:, V:].sum()) > 0
