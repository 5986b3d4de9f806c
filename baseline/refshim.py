"""Import the UNMODIFIED reference (``baseline/_ref`` snapshot, or /root/reference in the build container) and drive it
through its own public API.  Test / benchmark infrastructure only.

Import shim: ``bcolz`` / ``nltk`` are absent and ``from transformers import *`` (common/Utils.py:11, pinned
transformers 2.1.1) used to leak ``torch`` and ``math`` into the module; three stub modules restore exactly that.  No
reference file is modified.

Protocol adapters: ``Generations.beam`` / ``greedy`` (common/Generations.py:66-190) drive an ``EncDecModel``
(GTTP/EncDecModel.py:11-42).  The CaSE decoder never implemented that protocol, so ``CaseAdapter`` exposes it around the
unchanged ``CaSETransformerSeqDecoder`` (``generate`` = the module's own training branch with every dropout disabled,
i.e. the reference's math on a given prefix); ``make_gttp_adapter`` wraps GTTP's encode tuple in a dict so ``get_data``
can slice it.  All arithmetic stays inside reference code.
"""
import math
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_LOADED = {}


def reference_root():
    """baseline/_ref when the snapshot exists (GPU box), else /root/reference (build container), else None."""
    for p in (os.path.join(HERE, '_ref'), '/root/reference'):
        if os.path.isfile(os.path.join(p, 'CaSE', 'Model.py')):
            return p
    return None


def load_reference():
    """-> namespace with the reference modules (case, gttp, masque, glks, gen, utils, trainer).  Raises if the sources
    are not available."""
    if 'ns' in _LOADED:
        return _LOADED['ns']
    root = reference_root()
    if root is None:
        raise RuntimeError('reference sources not found: run `python baseline/make_ref.py` in the build container')
    for m in ('bcolz', 'nltk'):
        sys.modules.setdefault(m, types.ModuleType(m))
    tr = types.ModuleType('transformers')
    tr.torch, tr.math = torch, math
    tr.__all__ = ['torch', 'math']
    real_tr = sys.modules.get('transformers')
    sys.modules['transformers'] = tr
    sys.path.insert(0, root)
    try:
        import CaSE.Model as ref_case
        import GTTP.Model as ref_gttp
        import Masque.Model as ref_masque
        import GLKS.Model as ref_glks
        import common.Generations as ref_gen
        import common.Utils as ref_utils
    finally:
        if real_tr is not None:
            sys.modules['transformers'] = real_tr
        else:
            sys.modules.pop('transformers', None)
    ns = types.SimpleNamespace(root=root, case=ref_case, gttp=ref_gttp, masque=ref_masque, glks=ref_glks, gen=ref_gen,
                               utils=ref_utils)
    _LOADED['ns'] = ns
    return ns


class no_dropout:
    """Disable every dropout the decoder's training branch touches, without editing the reference."""

    def __init__(self, dec, ref_case):
        self.dec, self.ref_case = dec, ref_case

    def __enter__(self):
        self.saved = []
        for m in self.dec.modules():
            if isinstance(m, torch.nn.Dropout):
                self.saved.append((m, 'p', m.p)); m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention):
                self.saved.append((m, 'dropout', m.dropout)); m.dropout = 0.0
        self.fd = self.ref_case.F.dropout
        self.ref_case.F.dropout = lambda x, p=0.5, training=True, inplace=False: x
        self.dec.train()

    def __exit__(self, *a):
        self.ref_case.F.dropout = self.fd
        for m, k, v in self.saved:
            setattr(m, k, v)
        self.dec.eval()


def run_teacher_forced(ns, dec, mems, masks, weights, feat, onehot, prefix, BOS=1, UNK=100):
    """dist for every position of ``prefix`` ([R,n], BOS first) from the reference's own code (CaSE/Model.py:64-90)."""
    gt = torch.cat([prefix[:, 1:], torch.zeros(prefix.size(0), 1, dtype=torch.long, device=prefix.device)], 1)
    with torch.no_grad(), no_dropout(dec, ns.case):
        dec_out, gen, (d1, d2), _ = dec(mems, BOS, UNK, onehot, groundtruth_index=gt, additional_decoder_feature=feat,
                                        encode_weights=weights, encode_masks=masks)
    return d1 + d2, gen, dec_out


class CaseAdapter:
    """EncDecModel protocol over an unchanged CaSETransformerSeqDecoder."""

    def __init__(self, ns, dec, inp):
        self.ns, self.dec, self.inp = ns, dec, inp

    def encode(self, data):
        i = self.inp
        return {'mem_q': i.mem_q, 'mem_p': i.mem_p, 'mask_q': i.query.ne(0), 'mask_p': i.passage.ne(0),
                'w_q': i.prior_q, 'w_p': i.prior_p, 'feat': i.answer_rep}

    def init_decoder_states(self, data, enc):
        return torch.zeros(self.inp.query.size(0), 0, dtype=torch.long, device=self.inp.query.device)

    def generation_to_decoder_input(self, data, indices):
        return indices

    def decode(self, data, previous_word, enc, prev):
        return {'state': torch.cat([prev['state'], previous_word.view(-1, 1)], 1)}

    def generate(self, data, enc, dec_out, softmax=True):
        d, _, _ = run_teacher_forced(self.ns, self.dec, [enc['mem_q'], enc['mem_p']], [enc['mask_q'], enc['mask_p']],
                                     [enc['w_q'], enc['w_p']], enc['feat'], data['source_map'], dec_out['state'])
        return d[:, -1]

    def to_word(self, data, gen_output, k=5, sampling=False):
        return self.ns.utils.topk(gen_output, k=k)


def make_gttp_adapter(ns, *args, **kw):
    """GTTP with the bi-GRU encoders bypassed (their outputs are the inputs of the hot path) and the encode tuple
    wrapped in a dict so Generations.beam's get_data can slice it; decode / generate / to_word are the reference's."""

    class GttpAdapter(ns.gttp.GTTP):
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

    return GttpAdapter(*args, **kw)


def reference_decoder(ns, sd, V, H=256):
    dec = ns.case.CaSETransformerSeqDecoder(2, 4, 8, V, H)
    dec.load_state_dict(sd, strict=True)
    return dec.eval()


def reference_case_model(ns, V, T, decoder_sd=None, seed=1234, H=256):
    """A full reference CaSE model (encoders, passage selection, supporting-token identification, decoder) with
    xavier-initialised weights (common/CumulativeTrainer.py:13-24) and, optionally, a given decoder state."""
    from case_rg_b200 import synthetic as syn
    vocab2id, id2vocab = syn.make_vocab(V)
    torch.manual_seed(seed)
    full = ns.case.CaSE(4, T, id2vocab, vocab2id, H)
    for p in full.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p.data)
    if decoder_sd is not None:
        full.response_generation.decoder.load_state_dict(decoder_sd)
    return full.eval()
