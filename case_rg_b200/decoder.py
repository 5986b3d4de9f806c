"""Module face of the drop-in: ``FastCaSEDecoder`` stands where the reference builds
``CaSETransformerSeqDecoder`` (CaSE/Model.py:265) and keeps its constructor, parameter names
(so ``load_state_dict`` of a reference checkpoint works, CaSE/Run.py:55) and ``forward`` signature
/ 4-tuple return (CaSE/Model.py:50,125).

Differences, all in test mode only (training stays on the reference module):
  * ``source_map`` is taken in its int64 index form [B,S] (what CaSEDataset / collate_fn produce,
    CaSE/CaSEDataset.py:98-104,140).  The dense one-hot of ``build_map`` (Utils.py:344-355) is still
    accepted and converted back, but at BASELINE sizes it does not fit (20 GB at B=64, S=2620), so
    ``install_fast_decoder`` also stops ``CaSE.forward`` from building it.
  * the reference returns all T positions of its last step in elements [0..2]; its only test-mode consumer
    reads element [3] (CaSE/Model.py:331).  Here [0] covers the last decoded position only ([B,1,H]), [1] is None
    and [2] is None unless ``return_distribution`` is set (greedy only): the search path takes the top-k from the
    mixture's parts and never builds the [R,V] tensor, so asking for it costs one extra step.
  * ``beam_width`` (default 1 = the reference's in-module greedy loop) selects Generations.beam
    semantics on the device.
"""
import math
import types
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from .engine import CaseDecodeEngine, CaseWeights


class _PositionalTable(nn.Module):
    """Holds the ``pe`` buffer under the reference's key ``embedding.1.pe`` (PositionalEmbedding.py:27-32)."""

    def __init__(self, H, max_len=1000):
        super().__init__()
        pe = torch.zeros(max_len, H)
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, H, 2).float() * (-math.log(10000.0) / H))
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.register_buffer('pe', pe)


class _LayerParams(nn.Module):
    def __init__(self, H, nhead):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(H, nhead)
        self.multihead_attn = nn.MultiheadAttention(H, nhead)
        self.linear1, self.linear2 = nn.Linear(H, H), nn.Linear(H, H)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(H), nn.LayerNorm(H), nn.LayerNorm(H)


class _StackParams(nn.Module):
    def __init__(self, H, nhead, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([_LayerParams(H, nhead) for _ in range(num_layers)])


class _AdditiveParams(nn.Module):
    def __init__(self, q, k, h):
        super().__init__()
        self.linear_key = nn.Linear(k, h, bias=False)
        self.linear_query = nn.Linear(q, h, bias=True)
        self.v = nn.Linear(h, 1, bias=False)


class FastCaSEDecoder(nn.Module):
    """Parameter container with the reference's state_dict layout + the CUDA decode path."""

    def __init__(self, num_memories, num_layers, nhead, tgt_vocab_size, hidden_size, emb_matrix=None,
                 beam_width: int = 1, dtype: str = 'bf16', vocab_impl: Optional[int] = None, use_graph: bool = True,
                 opt: int = 0):
        super().__init__()
        self.opt = int(opt)                   # CASE_OPT_* bits of the engines (0 = default fast path)
        if (num_memories, num_layers, nhead, hidden_size) != (2, 4, 8, L.H):
            raise ValueError('FastCaSEDecoder is built for the shipped configuration: num_memories=2, '
                             f'num_layers=4, nhead=8, hidden_size={L.H} (CaSE/Model.py:265, Run.py:71)')
        H = hidden_size
        self.tgt_vocab_size, self.num_layers, self.hidden_size = tgt_vocab_size, num_layers, hidden_size
        self.embedding = nn.Sequential(nn.Embedding(tgt_vocab_size, H, padding_idx=0), _PositionalTable(H))
        if emb_matrix is not None:
            self.embedding[0].weight.data.copy_(torch.as_tensor(emb_matrix))
        self.decs = nn.ModuleList([_StackParams(H, nhead, num_layers) for _ in range(num_memories)])
        self.norm1, self.norm2 = nn.LayerNorm(H), nn.LayerNorm(H)
        self.attns = nn.ModuleList([_AdditiveParams(2 * H, H, H) for _ in range(num_memories)])
        self.gen = nn.Sequential(nn.Linear(3 * H, H), nn.Dropout(0.1), nn.Linear(H, tgt_vocab_size, bias=False),
                                 nn.Softmax(dim=-1))
        self.mix = nn.Linear(3 * H, num_memories + 1)
        self.beam_width, self.dtype, self.vocab_impl, self.use_graph = beam_width, dtype, vocab_impl, use_graph
        self.return_distribution = False      # True: forward()[2] is the last step's extended distribution (one extra step)
        self._weights = None
        self._engines: Dict[tuple, CaseDecodeEngine] = {}

    # ---------------------------------------------------------------- construction helpers
    @classmethod
    def from_reference(cls, ref_decoder: nn.Module, **kw) -> "FastCaSEDecoder":
        """Build from a reference ``CaSETransformerSeqDecoder`` instance (weights copied)."""
        sd = ref_decoder.state_dict()
        V, H = sd['embedding.0.weight'].shape
        m = cls(len(ref_decoder.decs), ref_decoder.num_layers, 8, V, H, **kw)
        m.load_state_dict(sd)
        return m.to(sd['embedding.0.weight'].device).eval()

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], **kw) -> "FastCaSEDecoder":
        V, H = sd['embedding.0.weight'].shape
        m = cls(2, 4, 8, V, H, **kw)
        m.load_state_dict(sd)
        return m.eval()

    def _load_from_state_dict(self, *a, **k):
        self._weights = None          # re-pack on next use
        self._engines = {}
        return super()._load_from_state_dict(*a, **k)

    def refresh_weights(self):
        self._weights, self._engines = None, {}

    def _packed(self, device) -> CaseWeights:
        if self._weights is None or self._weights.device != device or self._weights.dtype_name != self.dtype:
            self._weights = CaseWeights(self.state_dict(), device=device, dtype=self.dtype)
            self._engines = {}
        return self._weights

    def engine_for(self, B, W, S0, S1, T, device) -> CaseDecodeEngine:
        key = (B, W, S0, S1, T, str(device), self.dtype, self.vocab_impl, self.opt)
        e = self._engines.get(key)
        if e is None:
            if len(self._engines) >= 4:       # shapes are few in practice (full batches + one tail batch)
                self._engines.clear()
            e = CaseDecodeEngine(self._packed(device), B, W, S0, S1, T, vocab_impl=self.vocab_impl, opt=self.opt)
            self._engines[key] = e
        return e

    # ---------------------------------------------------------------- reference signature
    def forward(self, encode_memories, BOS, UNK, source_map, groundtruth_index=None, additional_decoder_feature=None,
                encode_weights=None, encode_masks=None, init_decoder_state=None, max_target_length=None):
        if self.training:
            raise NotImplementedError('FastCaSEDecoder covers the test-mode decode path only; train with the '
                                      'reference CaSETransformerSeqDecoder and load its checkpoint here')
        if BOS != 1:
            raise ValueError('BOS id must be 1 ([unused0], common/Constants.py:2)')
        dev = encode_memories[0].device
        if dev.type != 'cuda':
            raise RuntimeError('FastCaSEDecoder needs CUDA tensors: there is no CPU fallback')
        B = source_map.size(0)
        if source_map.dim() == 3:                 # dense one-hot from build_map: recover the indices
            source_map = source_map.argmax(dim=-1)
        if max_target_length is None:
            max_target_length = groundtruth_index.size(1)
        H = self.hidden_size
        mems = [m.reshape(B, -1, H) for m in encode_memories]
        W = int(self.beam_width)
        eng = self.engine_for(B, W, mems[0].size(1), mems[1].size(1), int(max_target_length), dev)
        eng.prefill(encode_memories[0], encode_memories[1], encode_masks[0], encode_masks[1], encode_weights[0],
                    encode_weights[1], additional_decoder_feature, source_map)
        mode = L.MODE_MODULE_GREEDY if W == 1 else L.MODE_BEAM
        tokens = eng.decode(int(max_target_length), mode, use_graph=self.use_graph)
        sel = slice(0, None, W)
        dec_outputs = eng.hN[sel].unsqueeze(1)
        ext = None                  # the search path never builds the [R,V] mixture (top-k is taken from its parts)
        if self.return_distribution and W == 1:
            # on request: the last step again, up to the finished distribution (idempotent for the in-module greedy loop)
            ext = eng.step_distribution(int(max_target_length) - 1)[:, :self.tgt_vocab_size].unsqueeze(1).clone()
        return dec_outputs, None, ext, tokens


def masque_to_case_state(msd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """State dict of ``MasqueTransformerSeqDecoder`` (Masque/Model.py:13-36) -> the CaSE decoder layout the engine packs.

    Masque's decoder is CaSE's without the additional decoder feature: one ``norm`` (= CaSE ``norm1``), attention
    queries from the H-wide decoder state only (``linear_query`` [H,H], Masque/Model.py:31 against CaSE/Model.py:32),
    ``gen`` = Linear(2H,H) -> Linear(H,V) (keys gen.0 / gen.1, no Dropout module in between).  The feature columns of
    CaSE's ``linear_query`` [H,2H] and ``gen.0`` [H,3H] are filled with exact zeros and the engine is fed a zero
    feature (norm2 = identity affine), so every product with them is exactly 0: same arithmetic, same kernels."""
    H = msd['norm.weight'].numel()
    sd = {}
    for k, v in msd.items():
        if k.startswith('norm.'):
            sd['norm1.' + k[5:]] = v
        elif k == 'gen.1.weight':
            sd['gen.2.weight'] = v
        elif k == 'gen.0.weight':
            sd[k] = torch.cat([v, v.new_zeros(v.size(0), H)], 1)
        elif k.endswith('linear_query.weight'):
            sd[k] = torch.cat([v, v.new_zeros(v.size(0), H)], 1)
        else:
            sd[k] = v
    sd['norm2.weight'] = msd['norm.weight'].new_ones(H)
    sd['norm2.bias'] = msd['norm.weight'].new_zeros(H)
    return sd


class FastMasqueDecoder(FastCaSEDecoder):
    """Module face for the Masque baseline: stands where the reference builds
    ``MasqueTransformerSeqDecoder`` (Masque/Model.py:13), keeps ITS state_dict keys and ``forward`` signature
    (Masque/Model.py:47: ``encode_masks`` / ``encode_weights`` before ``groundtruth_index``, no
    ``additional_decoder_feature``) and runs the CaSE decode kernels through ``masque_to_case_state``."""

    def __init__(self, num_memories, num_layers, nhead, tgt_vocab_size, hidden_size, emb_matrix=None, **kw):
        super().__init__(num_memories, num_layers, nhead, tgt_vocab_size, hidden_size, emb_matrix=emb_matrix, **kw)
        H = hidden_size
        del self.norm1, self.norm2
        self.norm = nn.LayerNorm(H)
        self.attns = nn.ModuleList([_AdditiveParams(H, H, H) for _ in range(num_memories)])
        self.gen = nn.Sequential(nn.Linear(2 * H, H), nn.Linear(H, tgt_vocab_size, bias=False), nn.Softmax(dim=-1))

    def _packed(self, device) -> CaseWeights:
        if self._weights is None or self._weights.device != device or self._weights.dtype_name != self.dtype:
            self._weights = CaseWeights(masque_to_case_state(self.state_dict()), device=device, dtype=self.dtype)
            self._engines = {}
        return self._weights

    def forward(self, encode_memories, BOS, UNK, source_map, encode_masks=None, encode_weights=None,
                groundtruth_index=None, init_decoder_state=None, max_target_length=None):
        B, dev = source_map.size(0), encode_memories[0].device
        if encode_weights is None:        # Masque/Model.py:100-103: without weights p is the attention itself
            encode_weights = [torch.ones(B, m.reshape(B, -1, self.hidden_size).size(1), device=dev) for m in encode_memories]
        feat = torch.zeros(B, self.hidden_size, device=dev)
        return super().forward(encode_memories, BOS, UNK, source_map, groundtruth_index=groundtruth_index,
                               additional_decoder_feature=feat, encode_weights=encode_weights,
                               encode_masks=encode_masks, init_decoder_state=init_decoder_state,
                               max_target_length=max_target_length)


def install_fast_decoder(model: nn.Module, beam_width: int = 1, dtype: str = 'bf16', **kw) -> nn.Module:
    """Swap the decoder of a reference ``CaSE`` model (CaSE/Model.py:255-339) - or of the ``Masque`` baseline
    (Masque/Model.py:202-285: same ``response_generation.decoder`` / ``do_test`` / ``build_map`` structure) - for the
    CUDA path, in place.

    * ``model.response_generation.decoder`` becomes a ``FastCaSEDecoder`` holding the same weights;
    * ``model.forward`` keeps ``data['source_map']`` in index form instead of calling ``build_map``
      (Model.py:334-335) when ``method == 'test'``; training goes through the original forward.
    """
    ref_dec = model.response_generation.decoder
    cls = FastMasqueDecoder if hasattr(ref_dec, 'norm') and not hasattr(ref_dec, 'norm1') else FastCaSEDecoder   # Masque/Model.py:29
    fast = cls.from_reference(ref_dec, beam_width=beam_width, dtype=dtype, **kw)
    model.response_generation.decoder = fast
    orig_forward = model.forward

    def forward(self, data, method='mle_train'):
        if method == 'test':
            return self.do_test(data)
        return orig_forward(data, method=method)

    model.forward = types.MethodType(forward, model)
    model._reference_decoder = [ref_dec]     # list: keeps it out of the module tree / state_dict
    return model


def install_fast_model(model: nn.Module, beam_width: int = 1, dtype: str = 'bf16', **kw) -> nn.Module:
    """The whole of ``CaSE.do_test`` on the device (CaSE/Model.py:313-331): ``install_fast_decoder`` plus the pre-decode
    producers - shared encoder, passage selection, supporting-token identification, prior / answer
    representation - through ``producers.CaseProducers`` built from the model's own state_dict.  ``model.forward(data,
    'test')`` keeps its contract ({'answer', 'rank'}); training goes through the original modules."""
    from .producers import CaseProducers
    install_fast_decoder(model, beam_width=beam_width, dtype=dtype, **kw)
    dev = next(model.parameters()).device
    prod = CaseProducers(model.state_dict(), device=dev)
    dec = model.response_generation.decoder

    def do_test(self, data):
        p = prod(data['query'], data['passage'])
        out = dec([p['mem_q'], p['mem_p']], self.response_generation.BOS, self.response_generation.UNK, data['source_map'],
                  additional_decoder_feature=p['answer_rep'], groundtruth_index=None, max_target_length=self.max_target_length,
                  encode_masks=[data['query'].ne(0), data['passage'].ne(0)], encode_weights=[p['prior_q'], p['prior_p']])
        return {'answer': out[3], 'rank': p['rank']}

    model.do_test = types.MethodType(do_test, model)
    model._fast_producers = [prod]
    return model


def install_fast_gttp(model: nn.Module, dtype: str = 'bf16', device=None, **kw) -> nn.Module:
    """Swap the step side of a reference ``GTTP`` model (GTTP/Model.py:133-212) for the CUDA path, in place.

    ``model.forward(data, 'test')`` keeps its contract (``{'answer': LongTensor}``, greedy for ``beam_width == 1``, else
    beam; GTTP/Model.py:204-212) but
      * ``data['background_map']`` stays in its int64 index form - ``build_map`` (Utils.py:344-355) is not called;
      * the bi-GRU encoders and ``enc2dec`` (``encode`` / ``init_decoder_states``, Model.py:156-174) run as the reference
        wrote them - they are outside the hot path - and their outputs feed the device search;
      * ``decode`` / ``generate`` / ``to_word`` + ``Generations.greedy/beam`` (Model.py:176-193, Generations.py:66-190) are
        replaced by ``FastGTTP.fast_search`` (gttp_decode_step x max_dec_len in one CUDA graph).
    Training (``method='train'``) goes through the original forward.

    ``device``: where the step engine lives (default: the model's device).  The reference's encoders call
    ``pack_padded_sequence`` with device-side lengths (common/Utils.py:313-336), which current torch only accepts on the
    CPU - so the model may stay on the CPU while the search runs on ``device``: the encoder outputs (a few MB) are moved
    there, the answers come back on the model's device."""
    from .generations import FastGTTP
    sd = {k: v for k, v in model.state_dict().items() if k.startswith('dec.') or k.startswith('gen.')}
    mdev = next(model.parameters()).device
    dev = torch.device(device) if device is not None else mdev
    fast = FastGTTP(sd, device=dev, dtype=dtype, max_dec_len=model.max_dec_len, beam_width=model.beam_width, **kw)
    orig_forward = model.forward

    def forward(self, data, method='train'):
        if method != 'test':
            return orig_forward(data, method=method)
        enc = self.encode(data)                                   # (c_enc_output, c_state, b_enc_output, b_state)
        init = self.init_decoder_states(data, enc)                # [B, 1, H]
        d = dict(context=data['context'], background=data['background'], background_map=data['background_map'],
                 src_output=enc[0], bg_output=enc[2], init_state=init)
        d = {k: v.to(dev) for k, v in d.items()}
        mode = L.MODE_PROTO_GREEDY if self.beam_width == 1 else L.MODE_BEAM
        return {'answer': fast.fast_search(d, self.max_dec_len, self.beam_width, mode).to(mdev)}

    model.forward = types.MethodType(forward, model)
    model._fast_gttp = [fast]            # list: keeps it out of the module tree / state_dict
    return model
