#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Runs only in the build container (it imports /root/reference, which does not exist on the GPU
box); the .npz files it writes are committed and are what the tests read.  The reference has no
tests or golden vectors of its own, so these outputs are the pin for ``oracle/``.

    python tests/golden/make_golden.py

Import shim: SURVEY.md §8c (bcolz / nltk absent; ``from transformers import *`` used to leak
``torch`` and ``math`` into common/Utils.py).  No reference file is modified or copied.

How the reference is driven
  * ``module_greedy``   : CaSETransformerSeqDecoder.forward in eval mode (CaSE/Model.py:91-123).
  * ``teacher_forced``  : the same module's training branch (Model.py:64-90) with every dropout
                          disabled, which evaluates the reference's own math on an arbitrary given
                          prefix (used for prefixes containing PAD, and as ``generate`` below).
  * ``beam`` / ``greedy``: common/Generations.py driven through a protocol adapter
                          (GTTP/EncDecModel.py:11-42) whose ``generate`` is the call above and
                          whose GTTP variant only wraps the encode tuple in a dict.
"""
import math
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def install_reference():
    for m in ('bcolz', 'nltk'):
        sys.modules[m] = types.ModuleType(m)
    tr = types.ModuleType('transformers')
    tr.torch, tr.math = torch, math
    tr.__all__ = ['torch', 'math']
    sys.modules['transformers'] = tr
    sys.path.insert(0, '/root/reference')


install_reference()
import CaSE.Model as ref_case            # noqa: E402
import GTTP.Model as ref_gttp            # noqa: E402
import common.Generations as ref_gen     # noqa: E402
from common.Utils import build_map, topk  # noqa: E402

from case_rg_b200 import synthetic as syn  # noqa: E402


def ref_decoder(sd, V, H=256):
    dec = ref_case.CaSETransformerSeqDecoder(2, 4, 8, V, H)
    dec.load_state_dict(sd, strict=True)
    return dec.eval()


def run_module_greedy(dec, inp, T):
    with torch.no_grad():
        out = dec(inp.encode_memories, syn.BOS, syn.UNK, build_map(inp.source_map, max=inp.V),
                  additional_decoder_feature=inp.answer_rep, encode_weights=inp.encode_weights,
                  encode_masks=inp.encode_masks, max_target_length=T)
    return out  # dec_outputs, gen_outputs, extended_gen_outputs, output_indexes


class no_dropout:
    """Disable every dropout the training branch touches, without editing the reference."""

    def __init__(self, dec):
        self.dec = dec

    def __enter__(self):
        self.saved = []
        for m in self.dec.modules():
            if isinstance(m, torch.nn.Dropout):
                self.saved.append((m, 'p', m.p)); m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention):
                self.saved.append((m, 'dropout', m.dropout)); m.dropout = 0.0
        self.fd = ref_case.F.dropout
        ref_case.F.dropout = lambda x, p=0.5, training=True, inplace=False: x
        self.dec.train()

    def __exit__(self, *a):
        ref_case.F.dropout = self.fd
        for m, k, v in self.saved:
            setattr(m, k, v)
        self.dec.eval()


def run_teacher_forced(dec, mems, masks, weights, feat, onehot, prefix):
    """dist for every position of ``prefix`` ([R,n], BOS first) from the reference's own code."""
    gt = torch.cat([prefix[:, 1:], torch.zeros(prefix.size(0), 1, dtype=torch.long)], 1)
    with torch.no_grad(), no_dropout(dec):
        dec_out, gen, (d1, d2), _ = dec(mems, syn.BOS, syn.UNK, onehot, groundtruth_index=gt,
                                        additional_decoder_feature=feat, encode_weights=weights,
                                        encode_masks=masks)
    return d1 + d2, gen, dec_out


class CaseAdapter:
    """EncDecModel protocol over an unchanged CaSETransformerSeqDecoder."""

    def __init__(self, dec, inp):
        self.dec, self.inp = dec, inp

    def encode(self, data):
        i = self.inp
        return {'mem_q': i.mem_q, 'mem_p': i.mem_p, 'mask_q': i.query.ne(0), 'mask_p': i.passage.ne(0),
                'w_q': i.prior_q, 'w_p': i.prior_p, 'feat': i.answer_rep}

    def init_decoder_states(self, data, enc):
        return torch.zeros(self.inp.query.size(0), 0, dtype=torch.long)

    def generation_to_decoder_input(self, data, indices):
        return indices

    def decode(self, data, previous_word, enc, prev):
        return {'state': torch.cat([prev['state'], previous_word.view(-1, 1)], 1)}

    def generate(self, data, enc, dec_out, softmax=True):
        d, _, _ = run_teacher_forced(self.dec, [enc['mem_q'], enc['mem_p']], [enc['mask_q'], enc['mask_p']],
                                     [enc['w_q'], enc['w_p']], enc['feat'], data['source_map'], dec_out['state'])
        return d[:, -1]

    def to_word(self, data, gen_output, k=5, sampling=False):
        return topk(gen_output, k=k)


class GttpAdapter(ref_gttp.GTTP):
    """GTTP with the bi-GRU encoders bypassed (their outputs are the fixture's inputs) and the
    encode tuple wrapped in a dict so Generations.beam's get_data can slice it."""

    def attach(self, inp):
        self._inp = inp

    def encode(self, data):
        return {'c': self._inp.src_output, 'b': self._inp.bg_output}

    def init_decoder_states(self, data, enc):
        return self._inp.init_state

    def decode(self, data, previous_word, enc, prev):
        feat, [st], [sa, ba], _ = self.dec(previous_word, prev['state'], enc['c'], enc['b'],
                                           src_mask=data['context'].ne(0), bg_mask=data['background'].ne(0))
        return {'state': st, 'feature': feat, 'bg_attn': ba}


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'wrote {path}  {os.path.getsize(path) / 1024:.0f} KiB')


def main():
    torch.set_num_threads(8)
    V, H = 1000, 256

    # ---- 1. module greedy, reference (xavier) init ---------------------------------------------
    cfg = dict(wseed=11, iseed=21, B=3, Lq=12, NP=3, Lp=16, T=8, peaked=0.0)
    sd = syn.make_case_decoder_state(cfg['wseed'], V, H)
    inp = syn.make_case_inputs(cfg['iseed'], cfg['B'], cfg['Lq'], cfg['NP'], cfg['Lp'], V, H)
    dec = ref_decoder(sd, V)
    dec_out, gen, dist, toks = run_module_greedy(dec, inp, cfg['T'])
    save('case_module_greedy_xavier', tokens=toks, dist=dist, gen=gen, dec_out=dec_out,
         wsum=syn.state_checksum(sd), cfg=np.array(list(cfg.values()), dtype=np.float64),
         cfg_keys=np.array(list(cfg.keys())))

    # ---- 2. module greedy, peaked init (robust argmax margins), longer ---------------------------
    cfg = dict(wseed=12, iseed=22, B=4, Lq=10, NP=4, Lp=12, T=12, peaked=0.35, pad_boost=14.0, gate=2.0)
    sd = syn.make_case_decoder_state(cfg['wseed'], V, H, peaked=cfg['peaked'], boost={0: cfg['pad_boost']},
                                     gen_gate_bias=cfg['gate'])
    inp = syn.make_case_inputs(cfg['iseed'], cfg['B'], cfg['Lq'], cfg['NP'], cfg['Lp'], V, H)
    dec = ref_decoder(sd, V)
    dec_out, gen, dist, toks = run_module_greedy(dec, inp, cfg['T'])
    print('peaked greedy tokens', toks.tolist())
    save('case_module_greedy_peaked', tokens=toks, dist=dist, gen=gen, dec_out=dec_out,
         wsum=syn.state_checksum(sd), cfg=np.array(list(cfg.values()), dtype=np.float64),
         cfg_keys=np.array(list(cfg.keys())))

    # ---- 3. teacher-forced prefixes containing PAD(0) -------------------------------------------
    g = torch.Generator().manual_seed(5)
    prefix = torch.randint(3, V, (cfg['B'], 7), generator=g)
    prefix[:, 0] = syn.BOS
    prefix[0, 3] = 0
    prefix[1, 5] = 0
    prefix[1, 6] = 0
    prefix[2, 1] = 0
    d, gen_tf, dec_tf = run_teacher_forced(dec, inp.encode_memories, inp.encode_masks, inp.encode_weights,
                                           inp.answer_rep, build_map(inp.source_map, max=V), prefix)
    assert torch.isfinite(d).all()
    save('case_teacher_forced_pad', prefix=prefix, dist=d, gen=gen_tf, dec_out=dec_tf,
         wsum=syn.state_checksum(sd), cfg=np.array(list(cfg.values()), dtype=np.float64),
         cfg_keys=np.array(list(cfg.keys())))

    # ---- 4. Generations.greedy / beam over CaSE (EOS boosted so hypotheses finish) ---------------
    cfg = dict(wseed=13, iseed=23, B=6, Lq=10, NP=3, Lp=12, T=9, peaked=0.25, eos_boost=13.0, gate=2.0)
    sd = syn.make_case_decoder_state(cfg['wseed'], V, H, peaked=cfg['peaked'], boost={syn.EOS: cfg['eos_boost']},
                                     gen_gate_bias=cfg['gate'])
    inp = syn.make_case_inputs(cfg['iseed'], cfg['B'], cfg['Lq'], cfg['NP'], cfg['Lp'], V, H)
    dec = ref_decoder(sd, V)
    vocab2id, id2vocab = syn.make_vocab(V)
    data = {'id': inp.ids, 'source_map': build_map(inp.source_map, max=V)}
    ad = CaseAdapter(dec, inp)
    out = {}
    out['greedy'] = ref_gen.greedy(ad, dict(data), vocab2id, cfg['T'])
    for w in (1, 2, 4, 8):
        out[f'beam{w}'] = ref_gen.beam(ad, dict(data), vocab2id, cfg['T'], w)
        print(f'case beam{w}', out[f'beam{w}'].tolist())
    print('case greedy', out['greedy'].tolist())
    save('case_generations', wsum=syn.state_checksum(sd), cfg=np.array(list(cfg.values()), dtype=np.float64),
         cfg_keys=np.array(list(cfg.keys())), **out)

    # ---- 5. GTTP greedy / beam + first-step distribution -----------------------------------------
    Vg = 1200
    cfg = dict(wseed=14, iseed=24, B=6, Lc=10, NP=3, Lp=12, T=9, peaked=0.3, eos_boost=5.0)
    sdg = syn.make_gttp_state(cfg['wseed'], Vg, H, H, peaked=cfg['peaked'], boost={syn.EOS: cfg['eos_boost']})
    ginp = syn.make_gttp_inputs(cfg['iseed'], cfg['B'], cfg['Lc'], cfg['NP'], cfg['Lp'], Vg, H)
    vocab2id, id2vocab = syn.make_vocab(Vg)
    torch.manual_seed(0)
    model = GttpAdapter(H, H, vocab2id, id2vocab, max_dec_len=cfg['T'])
    missing = model.load_state_dict(sdg, strict=False)
    assert not missing.unexpected_keys
    model.eval()
    model.attach(ginp)
    data = {'id': ginp.ids, 'context': ginp.context, 'background': ginp.background,
            'background_map': build_map(ginp.background_map, max=Vg)}
    out = {}
    with torch.no_grad():
        out['greedy'] = ref_gen.greedy(model, dict(data), vocab2id, cfg['T'])
        for w in (1, 4, 8):
            out[f'beam{w}'] = ref_gen.beam(model, dict(data), vocab2id, cfg['T'], w)
            print(f'gttp beam{w}', out[f'beam{w}'].tolist())
        print('gttp greedy', out['greedy'].tolist())
        enc = model.encode(data)
        d0 = model.decode(data, torch.full((cfg['B'],), syn.BOS, dtype=torch.long), enc, {'state': ginp.init_state})
        dist0 = model.generate(data, enc, d0)
    save('gttp_generations', wsum=syn.state_checksum(sdg), cfg=np.array(list(cfg.values()), dtype=np.float64),
         cfg_keys=np.array(list(cfg.keys())), dist0=dist0, feat0=d0['feature'], state0=d0['state'],
         bg_attn0=d0['bg_attn'], **out)

    # ---- 6. full CaSE.forward(data,'test'): capture what the decoder is handed (boundary a15) ------
    cfg = dict(wseed=15, iseed=25, B=2, Lq=10, NP=3, Lp=12, T=6)
    vocab2id, id2vocab = syn.make_vocab(V)
    torch.manual_seed(1234)
    full = ref_case.CaSE(4, cfg['T'], id2vocab, vocab2id, H)
    for p in full.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p.data)
    sd = syn.make_case_decoder_state(cfg['wseed'], V, H, peaked=0.3, gen_gate_bias=2.0)
    full.response_generation.decoder.load_state_dict(sd)
    full.eval()
    inp = syn.make_case_inputs(cfg['iseed'], cfg['B'], cfg['Lq'], cfg['NP'], cfg['Lp'], V, H)
    captured = {}
    orig = full.response_generation.decoder.forward

    def spy(encode_memories, BOS, UNK, source_map, **kw):
        captured.update(mem_q=encode_memories[0], mem_p=encode_memories[1], feat=kw['additional_decoder_feature'],
                        w_q=kw['encode_weights'][0], w_p=kw['encode_weights'][1],
                        mask_q=kw['encode_masks'][0], mask_p=kw['encode_masks'][1],
                        max_target_length=kw['max_target_length'], BOS=BOS, UNK=UNK)
        return orig(encode_memories, BOS, UNK, source_map, **kw)

    full.response_generation.decoder.forward = spy
    with torch.no_grad():
        res = full({'id': inp.ids, 'query': inp.query, 'passage': inp.passage, 'source_map': inp.source_map.clone()},
                   method='test')
    save('case_model_forward_capture', answer=res['answer'], rank=res['rank'], query=inp.query, passage=inp.passage,
         source_map=inp.source_map, wsum=syn.state_checksum(sd),
         cfg=np.array(list(cfg.values()), dtype=np.float64), cfg_keys=np.array(list(cfg.keys())),
         **{k: (v if isinstance(v, torch.Tensor) else np.array(v)) for k, v in captured.items()})


if __name__ == '__main__':
    main()
