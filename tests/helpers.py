"""Shared fixture plumbing: rebuild the seeded weights / inputs a golden .npz was generated from."""
import os

import numpy as np
import torch

from case_rg_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
V_SMALL, V_GTTP_SMALL, H = 1000, 1200, 256


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    cfg = dict(zip([str(k) for k in z['cfg_keys']], [float(v) for v in z['cfg']]))
    return z, cfg


def case_state_for(name, cfg):
    w = int(cfg['wseed'])
    if name == 'case_module_greedy_xavier':
        return syn.make_case_decoder_state(w, V_SMALL, H)
    if name in ('case_module_greedy_peaked', 'case_teacher_forced_pad'):
        return syn.make_case_decoder_state(w, V_SMALL, H, peaked=cfg['peaked'], boost={0: cfg['pad_boost']},
                                           gen_gate_bias=cfg['gate'])
    if name == 'case_generations':
        return syn.make_case_decoder_state(w, V_SMALL, H, peaked=cfg['peaked'], boost={syn.EOS: cfg['eos_boost']},
                                           gen_gate_bias=cfg['gate'])
    if name == 'case_model_forward_capture':
        return syn.make_case_decoder_state(w, V_SMALL, H, peaked=0.3, gen_gate_bias=2.0)
    raise KeyError(name)


def build_case(name):
    """-> (npz, cfg, state_dict, CaseInputs) with the weight checksum verified."""
    z, cfg = load_golden(name)
    sd = case_state_for(name, cfg)
    assert abs(syn.state_checksum(sd) - float(z['wsum'])) < 1e-6 * max(1.0, abs(float(z['wsum']))), \
        'seeded weights differ from the ones the golden was generated with (torch RNG drift?)'
    inp = syn.make_case_inputs(int(cfg['iseed']), int(cfg['B']), int(cfg['Lq']), int(cfg['NP']), int(cfg['Lp']),
                               V_SMALL, H)
    return z, cfg, sd, inp


def build_gttp(name='gttp_generations'):
    z, cfg = load_golden(name)
    sd = syn.make_gttp_state(int(cfg['wseed']), V_GTTP_SMALL, H, H, peaked=cfg['peaked'],
                             boost={syn.EOS: cfg['eos_boost']})
    assert abs(syn.state_checksum(sd) - float(z['wsum'])) < 1e-6 * max(1.0, abs(float(z['wsum'])))
    inp = syn.make_gttp_inputs(int(cfg['iseed']), int(cfg['B']), int(cfg['Lc']), int(cfg['NP']), int(cfg['Lp']),
                               V_GTTP_SMALL, H)
    return z, cfg, sd, inp


def captured_inputs(z):
    """CaseInputs rebuilt from what the reference's CaSE.forward handed to its decoder."""
    t = lambda k: torch.from_numpy(z[k])
    B = z['query'].shape[0]
    return syn.CaseInputs(t('query'), t('passage'), t('source_map'), t('mem_q'), t('mem_p'), t('w_q'), t('w_p'),
                          t('feat'), torch.arange(B), V_SMALL)


def pad_to(tokens, L):
    out = torch.zeros(tokens.size(0), L, dtype=torch.long)
    out[:, :tokens.size(1)] = tokens
    return out


def build_masque(name='masque_module_greedy'):
    """-> (npz, cfg, Masque state_dict, CaseInputs) for the Masque decoder fixture (tests/golden/make_masque_golden.py)."""
    z, cfg = load_golden(name)
    msd = syn.make_masque_decoder_state(int(cfg['wseed']), V_SMALL, H, peaked=cfg['peaked'], boost={0: cfg['pad_boost']},
                                        gen_gate_bias=cfg['gate'])
    assert abs(syn.state_checksum(msd) - float(z['wsum'])) < 1e-6 * max(1.0, abs(float(z['wsum'])))
    inp = syn.make_case_inputs(int(cfg['iseed']), int(cfg['B']), int(cfg['Lq']), int(cfg['NP']), int(cfg['Lp']),
                               V_SMALL, H)
    return z, cfg, msd, inp


def glks_inputs(seed=31, R=6, V=700, Lb=90, H=256, D=40):
    """Seeded inputs of the GLKS Mixturer / copy_topk fixture (tests/golden/make_glks_golden.py reads them from here)."""
    g = torch.Generator().manual_seed(seed)
    state = torch.randn(R, 1, H, generator=g)
    p_v = torch.softmax(torch.randn(R, V, generator=g) * 2, 1)
    p_k = torch.softmax(torch.randn(R, Lb, generator=g) * 2, 1)
    p_k[:, ::9] = 0.0                                             # masked background positions carry no mass
    bmap = torch.randint(0, V, (R, Lb), generator=g)
    bmap[:, 5] = bmap[:, 6] = bmap[:, 7]                          # repeated copy targets
    w = torch.randn(1, H, generator=g) * 0.2
    b = torch.randn(1, generator=g)
    # copy_topk: extended rows [V + D], D dynamic words, some of them vocabulary words (overlap 0)
    gen_ext = torch.softmax(torch.randn(R, V + D, generator=g) * 2, 1)
    vmap = torch.randint(0, V, (R, D), generator=g)
    vmap[:, 3] = vmap[:, 4] = 100                                 # two OOV words fold onto UNK
    overlap = (torch.rand(R, D, generator=g) > 0.5).float()
    overlap[:, 3] = overlap[:, 4] = 1.0
    return dict(state=state, p_v=p_v, p_k=p_k, bmap=bmap, w=w, b=b, gen_ext=gen_ext, vmap=vmap, overlap=overlap)


def producers_case(ns=None, V=600, T=5, B=2, Lq=12, NP=3, Lp=20, seed=77):
    """Seeded full-model fixture of the pre-decode producers: (cfg, producer state_dict, decoder state_dict, data, model).
    The weights come from ``synthetic`` (no reference needed at test time); with ``ns`` (the loaded reference) the
    unmodified ``CaSE`` model is also built and loaded with exactly these weights (tests/golden/make_producers_golden.py,
    live drop-in tests)."""
    cfg = dict(V=V, T=T, B=B, Lq=Lq, NP=NP, Lp=Lp, seed=seed)
    inp = syn.make_case_inputs(seed + 1, B, Lq, NP, Lp, V, H)
    data = {'id': inp.ids, 'query': inp.query, 'passage': inp.passage, 'source_map': inp.source_map}
    sd_prod = syn.make_case_producer_state(seed, V, H)
    sd_dec = syn.make_case_decoder_state(seed + 2, V, H, peaked=0.3, gen_gate_bias=2.0)
    model = None
    if ns is not None:
        from baseline import refshim
        model = refshim.reference_case_model(ns, V, T, decoder_sd=sd_dec, seed=seed)
        missing = model.load_state_dict(sd_prod, strict=False)
        assert not missing.unexpected_keys, missing.unexpected_keys
        model.eval()
    return cfg, sd_prod, sd_dec, data, model
