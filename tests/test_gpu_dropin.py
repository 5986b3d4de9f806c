"""The drop-in boundary exercised LIVE on the device: the unmodified reference model
(snapshot under baseline/_ref, made by baseline/make_ref.py) runs ``model(data, method='test')`` on cuda:0, then
``install_fast_decoder`` / ``install_fast_gttp`` swap this repository's CUDA path into THE SAME model object and the same
call must return the same answers (fp32 storage: token for token; bf16: through the tie criterion of parity_tools).

Reference call chain: CaSE/Run.py:54-62 -> CumulativeTrainer.predict -> CaSE.forward (CaSE/Model.py:333-339) -> do_test
(:313-331) -> ResponseGeneration.action (:230-253) -> decoder; Masque/Model.py:266-285; GTTP/Model.py:204-212.
"""
import copy

import pytest
import torch

from case_rg_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
H = 256


@pytest.fixture(scope='module')
def ns():
    from baseline import refshim
    if refshim.reference_root() is None:
        pytest.skip('reference snapshot baseline/_ref not present (python baseline/make_ref.py)')
    return refshim.load_reference()


def _case_batch(seed, B, Lq, NP, Lp, V):
    inp = syn.make_case_inputs(seed, B, Lq, NP, Lp, V, H)
    return {'id': inp.ids.to(DEV), 'query': inp.query.to(DEV), 'passage': inp.passage.to(DEV),
            'source_map': inp.source_map.to(DEV)}


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_install_fast_decoder_into_live_reference_case(ns, dtype):
    """Whole CaSE.forward(data, 'test'): reference encoders / passage selection / supporting-token identification feed
    first the reference decoder, then the installed FastCaSEDecoder - on the same device, same weights, same batch."""
    from baseline import refshim
    from case_rg_b200.decoder import install_fast_decoder, FastCaSEDecoder
    V, T, B = 2000, 10, 4
    sd = syn.make_case_decoder_state(91, V, H, peaked=0.3, gen_gate_bias=2.0)
    model = refshim.reference_case_model(ns, V, T, decoder_sd=sd).to(DEV).eval()
    data = _case_batch(92, B, 20, 3, 24, V)
    with torch.no_grad():
        want = model(copy.copy(data), method='test')            # the unmodified model, one-hot source_map and all
        install_fast_decoder(model, dtype=dtype)
        assert isinstance(model.response_generation.decoder, FastCaSEDecoder)
        got = model(copy.copy(data), method='test')
    assert got['answer'].dtype == want['answer'].dtype == torch.int64
    assert got['answer'].shape == want['answer'].shape == (B, T)
    assert torch.equal(got['rank'], want['rank'])                 # the producers are untouched
    if dtype == 'fp32':
        assert torch.equal(got['answer'], want['answer']), (got['answer'], want['answer'])
    else:
        assert float((got['answer'] == want['answer']).float().mean()) >= 0.9, (got['answer'], want['answer'])
    # the state_dict of the patched model is still the reference's: a checkpoint round-trips (Run.py:55)
    ref_keys = set(refshim.reference_case_model(ns, V, T).state_dict().keys())
    assert set(model.state_dict().keys()) == ref_keys


def test_install_fast_model_into_live_reference_case(ns):
    """install_fast_model: producers AND decoder of the live reference CaSE replaced - forward(data, 'test') from token
    ids.  rank within the bf16 band of the unmodified model's; answers identical or oracle-rated ties (two bf16 stages)."""
    import parity_tools as PT
    from helpers import producers_case
    from case_rg_b200.decoder import install_fast_model
    from oracle.case_decoder import CaseOracle
    from oracle.producers import producers
    cfg, sd_prod, sd_dec, data, model = producers_case(ns)
    model = model.to(DEV).eval()
    d = {k: v.to(DEV) for k, v in data.items()}
    with torch.no_grad():
        want = model(copy.copy(d), method='test')
        install_fast_model(model, dtype='bf16')
        got = model(copy.copy(d), method='test')
    assert got['answer'].shape == want['answer'].shape and got['rank'].shape == want['rank'].shape
    assert float((got['rank'] - want['rank']).abs().max() / want['rank'].abs().max()) < 5e-2
    o = producers(sd_prod, data['query'], data['passage'])
    inp = syn.CaseInputs(data['query'], data['passage'], data['source_map'], o['mem_q'], o['mem_p'], o['prior_q'], o['prior_p'],
                         o['answer_rep'], data['id'], cfg['V'])
    orc = CaseOracle(sd_dec)
    ref, dists, scale = PT.oracle_greedy(lambda: orc.incremental(inp), cfg['B'], cfg['T'])
    assert torch.equal(ref, want['answer'].cpu())                  # oracle == the unmodified model on the device
    res = PT.compare_greedy(got['answer'].cpu(), ref, dists, 2 * 2e-2 * scale)
    assert res['miss'] == 0, res


def test_install_fast_decoder_into_live_reference_masque(ns):
    from case_rg_b200.decoder import install_fast_decoder, FastMasqueDecoder
    V, T, B = 2000, 8, 3
    vocab2id, id2vocab = syn.make_vocab(V)
    torch.manual_seed(7)
    model = ns.masque.Masque(T, id2vocab, vocab2id, H)
    for p in model.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p.data)
    msd = syn.make_masque_decoder_state(93, V, H, peaked=0.3, gen_gate_bias=2.0)
    model.response_generation.decoder.load_state_dict(msd)
    model = model.to(DEV).eval()
    data = _case_batch(94, B, 20, 3, 24, V)
    with torch.no_grad():
        want = model(copy.copy(data), method='test')
        install_fast_decoder(model, dtype='fp32')
        assert isinstance(model.response_generation.decoder, FastMasqueDecoder)
        got = model(copy.copy(data), method='test')
    assert torch.equal(got['answer'], want['answer']), (got['answer'], want['answer'])


@pytest.mark.parametrize('width', [1, 4])
def test_install_fast_gttp_into_live_reference_gttp(ns, width):
    """GTTP.forward(data, 'test'): the reference bi-GRU encoders feed first the reference decode/generate/to_word under
    Generations.greedy / beam (as shipped; beam through the dict-wrapping adapter of SURVEY.md §8c), then the installed
    device search.  The reference model stays on the CPU: its encoders pass device-side lengths to
    pack_padded_sequence (common/Utils.py:313-336), which this torch rejects on CUDA - the step engine runs on cuda:0."""
    from baseline import refshim
    from case_rg_b200.decoder import install_fast_gttp
    V, T, B = 1500, 9, 4
    vocab2id, id2vocab = syn.make_vocab(V)
    torch.manual_seed(11)
    model = ns.gttp.GTTP(H, H, vocab2id, id2vocab, max_dec_len=T, beam_width=width)
    sdg = syn.make_gttp_state(95, V, H, H, peaked=0.3, boost={syn.EOS: 5.0})
    missing = model.load_state_dict(sdg, strict=False)
    assert not missing.unexpected_keys
    model = model.eval()
    ginp = syn.make_gttp_inputs(96, B, 12, 3, 16, V, H)
    data = {'id': ginp.ids, 'context': ginp.context, 'background': ginp.background, 'background_map': ginp.background_map}
    # the reference helpers jump to CUDA whenever it is visible (Utils.new_tensor / build_map): pin them to the CPU for
    # the reference half of this test, the way bench.py's reference arm does with CUDA_VISIBLE_DEVICES=""
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        with torch.no_grad():
            if width == 1:
                want = model(copy.copy(data), method='test')['answer']      # the shipped path: Generations.greedy
            else:
                # Generations.beam slices encode outputs per node (Utils.get_data), which needs them in a dict: drive the
                # reference's own decode / generate / to_word through the adapter subclass
                ad = refshim.make_gttp_adapter(ns, H, H, vocab2id, id2vocab, max_dec_len=T, beam_width=width)
                ad.load_state_dict(model.state_dict())
                ad = ad.eval()
                enc = model.encode(data)
                fake = type('I', (), {})()
                fake.src_output, fake.bg_output, fake.init_state = enc[0], enc[2], model.init_decoder_states(data, enc)
                ad.attach(fake)
                d2 = dict(data, background_map=ns.utils.build_map(data['background_map'], max=V))
                want = ns.gen.beam(ad, d2, vocab2id, T, width)
    finally:
        torch.cuda.is_available = real
    with torch.no_grad():
        install_fast_gttp(model, dtype='fp32', device=DEV)
        got = model(copy.copy(data), method='test')['answer']
    assert got.dtype == torch.int64 and got.device == want.device
    L = min(got.size(1), want.size(1))
    assert torch.equal(got[:, :L], want[:, :L]), (got, want)
    assert got.size(1) == want.size(1) or width == 1
