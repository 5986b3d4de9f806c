"""Seeded synthetic CAsT-shaped inputs and random-init checkpoints (CPU, torch only).

There is no network for datasets or checkpoints, so every test and benchmark uses
tensors produced here.  Shapes and special ids mirror the reference:

* token layout of ``query`` / ``passage`` / ``source_map``: CaSE/CaSEDataset.py:59-106,
  collate at CaSE/CaSEDataset.py:130-140 (``source_map`` = ids of query ++ passages)
* GTTP ``context`` / ``background`` / ``background_map``: GTTP/GTTPDataset.py:40-95
* special ids (BERT-uncased): common/Constants.py:1-7
* checkpoint key names / shapes: ``CaSETransformerSeqDecoder.__init__`` CaSE/Model.py:14-36,
  ``BBCDecoder``/``CopyGenerator`` GTTP/Model.py:5-12,96-111
* init: xavier-uniform on every >1-d tensor, as ``init_params`` common/CumulativeTrainer.py:13-24

Nothing here touches CUDA; callers move tensors where they need them.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

PAD, BOS, EOS, UNK, CLS, SEP = 0, 1, 2, 100, 101, 102
BERT_VOCAB = 30522


# --------------------------------------------------------------------------- weights
def _xavier(gen, *shape):
    fan_out, fan_in = shape[0], shape[1]
    a = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(*shape, generator=gen) * 2 - 1) * a


def _uni(gen, n, a):
    return (torch.rand(n, generator=gen) * 2 - 1) * a


def sinusoid_table(max_len: int, H: int) -> torch.Tensor:
    """The ``pe`` buffer of common/PositionalEmbedding.py:27-32 (persisted in the state_dict)."""
    pe = torch.zeros(max_len, H)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, H, 2).float() * (-math.log(10000.0) / H))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def make_masque_decoder_state(seed: int, V: int = BERT_VOCAB, H: int = 256, **kw) -> Dict[str, torch.Tensor]:
    """Random-init state_dict with the keys of ``MasqueTransformerSeqDecoder`` (Masque/Model.py:14-36): the seeded CaSE
    state with the feature columns dropped (``linear_query`` [H,H], ``gen.0`` [H,2H]), ``norm1`` as ``norm`` and the
    vocabulary projection under ``gen.1``."""
    sd = make_case_decoder_state(seed, V, H, **kw)
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.startswith('norm2.'):
            continue
        if k.startswith('norm1.'):
            out['norm.' + k[6:]] = v
        elif k == 'gen.2.weight':
            out['gen.1.weight'] = v
        elif k == 'gen.0.weight':
            out[k] = v[:, :2 * H].contiguous()
        elif k.endswith('linear_query.weight'):
            out[k] = v[:, :H].contiguous()
        else:
            out[k] = v
    return out


def make_case_decoder_state(seed: int, V: int = BERT_VOCAB, H: int = 256, num_memories: int = 2,
                            num_layers: int = 4, peaked: float = 0.0,
                            boost: Optional[Dict[int, float]] = None,
                            gen_gate_bias: float = 0.0) -> Dict[str, torch.Tensor]:
    """Random-init state_dict with the 163 keys of ``CaSETransformerSeqDecoder`` (CaSE/Model.py:14-36).

    ``peaked`` > 0 scales ``gen.2.weight`` so the vocabulary softmax is far from uniform (argmax
    margins well above fp32 noise); 0 keeps the reference's xavier init.  ``boost`` maps a token id
    to an additive logit shift (used to make EOS / PAD show up in parity cases); ``gen_gate_bias``
    is added to the vocabulary component of the mixture gate so copying does not always win.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    sd['embedding.0.weight'] = _xavier(g, V, H)
    sd['embedding.1.pe'] = sinusoid_table(1000, H)
    for i in range(num_memories):
        for l in range(num_layers):
            p = f'decs.{i}.layers.{l}.'
            for att in ('self_attn', 'multihead_attn'):
                sd[p + att + '.in_proj_weight'] = _xavier(g, 3 * H, H)
                sd[p + att + '.in_proj_bias'] = _uni(g, 3 * H, 0.05)
                sd[p + att + '.out_proj.weight'] = _xavier(g, H, H)
                sd[p + att + '.out_proj.bias'] = _uni(g, H, 0.05)
            for lin in ('linear1', 'linear2'):
                sd[p + lin + '.weight'] = _xavier(g, H, H)
                sd[p + lin + '.bias'] = _uni(g, H, 1.0 / math.sqrt(H))
            for n in ('norm1', 'norm2', 'norm3'):
                sd[p + n + '.weight'] = 1.0 + _uni(g, H, 0.1)
                sd[p + n + '.bias'] = _uni(g, H, 0.1)
    for n in ('norm1', 'norm2'):
        sd[n + '.weight'] = 1.0 + _uni(g, H, 0.1)
        sd[n + '.bias'] = _uni(g, H, 0.1)
    for i in range(num_memories):
        sd[f'attns.{i}.linear_key.weight'] = _xavier(g, H, H)
        sd[f'attns.{i}.linear_query.weight'] = _xavier(g, H, 2 * H)
        sd[f'attns.{i}.linear_query.bias'] = _uni(g, H, 1.0 / math.sqrt(2 * H))
        sd[f'attns.{i}.v.weight'] = _xavier(g, 1, H)
    sd['gen.0.weight'] = _xavier(g, H, 3 * H)
    sd['gen.0.bias'] = _uni(g, H, 1.0 / math.sqrt(3 * H))
    sd['gen.2.weight'] = _xavier(g, V, H)
    sd['mix.weight'] = _xavier(g, num_memories + 1, 3 * H)
    sd['mix.bias'] = _uni(g, num_memories + 1, 1.0 / math.sqrt(3 * H))
    if peaked > 0:
        sd['gen.2.weight'] = torch.randn(V, H, generator=g) * peaked
        sd['attns.0.v.weight'] = sd['attns.0.v.weight'] * 4
        sd['attns.1.v.weight'] = sd['attns.1.v.weight'] * 4
    if boost:
        # shift the logit of ``tok`` by ~``s`` for every row: gen.0.bias gets 2*d (|d| = 1) and the
        # token's output row gets (s/2)*d, so the added logit is s plus a small input-dependent part.
        d = torch.randn(H, generator=g)
        d = d / d.norm()
        sd['gen.0.bias'] = sd['gen.0.bias'] + 2.0 * d
        for tok, s in boost.items():
            sd['gen.2.weight'][tok] = sd['gen.2.weight'][tok] + 0.5 * s * d
    if gen_gate_bias:
        sd['mix.bias'][0] = sd['mix.bias'][0] + gen_gate_bias
    return sd


def make_case_producer_state(seed: int, V: int = BERT_VOCAB, H: int = 256) -> Dict[str, torch.Tensor]:
    """Random-init state of CaSE's pre-decode producers under the reference's canonical key names (CaSE/Model.py:255-268):
    the shared ``query_encoder`` (TransformerSeqEncoder: 3 layers, 8 heads), ``passage_selection`` (Interaction, query
    blocks 5H->H + 2 x H->H, passage blocks 5H->H + 4 x H->H, scorer) and ``span_extraction`` (Interaction, query blocks
    5H->H + 1, passage blocks 5H->H + 2, norm1 / norm2, scorer).  xavier-uniform on matrices as ``init_params``
    (common/CumulativeTrainer.py:13-24); LayerNorm parameters and biases get small perturbations so that every term of
    the computation is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def mha(p, C):
        sd[p + 'in_proj_weight'] = _xavier(g, 3 * C, C)
        sd[p + 'in_proj_bias'] = _uni(g, 3 * C, 0.05)
        sd[p + 'out_proj.weight'] = _xavier(g, C, C)
        sd[p + 'out_proj.bias'] = _uni(g, C, 0.05)

    def norm(p, C):
        sd[p + 'weight'] = 1.0 + _uni(g, C, 0.1)
        sd[p + 'bias'] = _uni(g, C, 0.1)

    def linear(p, n, k):
        sd[p + 'weight'] = _xavier(g, n, k)
        sd[p + 'bias'] = _uni(g, n, 1.0 / math.sqrt(k))

    sd['query_encoder.embedding.0.weight'] = _xavier(g, V, H)
    sd['query_encoder.embedding.0.weight'][0] = 0.0                      # padding_idx = 0
    sd['query_encoder.embedding.1.pe'] = sinusoid_table(1000, H)
    for l in range(3):
        p = f'query_encoder.enc.layers.{l}.'
        mha(p + 'self_attn.', H)
        linear(p + 'linear1.', H, H)
        linear(p + 'linear2.', H, H)
        norm(p + 'norm1.', H)
        norm(p + 'norm2.', H)

    def blocks(prefix, n_extra):
        for i in range(1 + n_extra):
            C = 5 * H if i == 0 else H
            p = f'{prefix}{i}.'
            mha(p + 'self_attn.', C)
            norm(p + 'norm1.', C)
            norm(p + 'norm2.', C)
            linear(p + 'linear1.', H, C)
            linear(p + 'linear2.', H, H)

    for mod, nq, npb in (('passage_selection.', 2, 4), ('span_extraction.', 1, 2)):
        sd[mod + 'interaction.dual_att_linear.weight'] = _xavier(g, 1, 3 * H)
        blocks(mod + 'query_blocks.', nq)
        blocks(mod + 'passage_blocks.', npb)
        linear(mod + 'scorer.', 1, H)
    norm('span_extraction.norm1.', H)
    norm('span_extraction.norm2.', H)
    return sd


def make_gttp_state(seed: int, V: int = 50000, H: int = 256, E: int = 256, peaked: float = 0.0,
                    boost: Optional[Dict[int, float]] = None) -> Dict[str, torch.Tensor]:
    """Random-init step-side keys of GTTP (``dec.*`` and ``gen.*``; GTTP/Model.py:5-12,96-111)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    sd['dec.embedding.weight'] = _xavier(g, V, E)
    for a in ('src_attn', 'bg_attn'):
        sd[f'dec.{a}.linear_key.weight'] = _xavier(g, H, 2 * H)
        sd[f'dec.{a}.linear_query.weight'] = _xavier(g, H, H)
        sd[f'dec.{a}.linear_query.bias'] = _uni(g, H, 1.0 / math.sqrt(H))
        sd[f'dec.{a}.v.weight'] = _xavier(g, 1, H)
    sd['dec.gru.weight_ih_l0'] = _xavier(g, 3 * H, 4 * H + E)
    sd['dec.gru.weight_hh_l0'] = _xavier(g, 3 * H, H)
    sd['dec.gru.bias_ih_l0'] = _uni(g, 3 * H, 1.0 / math.sqrt(H))
    sd['dec.gru.bias_hh_l0'] = _uni(g, 3 * H, 1.0 / math.sqrt(H))
    sd['dec.readout.weight'] = _xavier(g, H, E + 5 * H)
    sd['dec.readout.bias'] = _uni(g, H, 1.0 / math.sqrt(E + 5 * H))
    sd['gen.linear.weight'] = _xavier(g, V, H)
    sd['gen.linear.bias'] = _uni(g, V, 1.0 / math.sqrt(H))
    sd['gen.linear_copy.weight'] = _xavier(g, 1, H)
    sd['gen.linear_copy.bias'] = _uni(g, 1, 1.0 / math.sqrt(H))
    if peaked > 0:
        sd['gen.linear.weight'] = torch.randn(V, H, generator=g) * peaked
        sd['dec.src_attn.v.weight'] = sd['dec.src_attn.v.weight'] * 8
        sd['dec.bg_attn.v.weight'] = sd['dec.bg_attn.v.weight'] * 8
    if boost:
        for tok, s in boost.items():
            sd['gen.linear.bias'][tok] = sd['gen.linear.bias'][tok] + s
    return sd


def state_checksum(sd: Dict[str, torch.Tensor]) -> float:
    """Order-independent fingerprint used by the golden fixtures to detect RNG drift."""
    tot = 0.0
    for k in sorted(sd):
        t = sd[k].double()
        tot += float(t.sum()) + 0.5 * float((t * t).sum())
    return tot


# --------------------------------------------------------------------------- inputs
def _token_rows(gen, n_rows: int, L: int, V: int, lo_frac: float, empty_frac: float, with_sep_mid: bool):
    """[CLS] tokens.. [SEP] PAD.. rows (CaSEDataset.py:77-87); ``with_sep_mid`` = the query form
    [CLS] ctx.. [SEP] query.. PAD.. (CaSEDataset.py:67-72)."""
    ids = torch.zeros(n_rows, L, dtype=torch.long)
    lo_tok = min(1000, max(V // 2, 110))
    for r in range(n_rows):
        if float(torch.rand((), generator=gen)) < empty_frac:
            ids[r, 0], ids[r, 1] = CLS, SEP
            continue
        n = int(torch.randint(max(3, int(lo_frac * L)), L + 1, (), generator=gen))
        body = torch.randint(lo_tok, V, (n,), generator=gen)
        body[0] = CLS
        if with_sep_mid:
            body[max(1, n // 2)] = SEP
        else:
            body[n - 1] = SEP
        ids[r, :n] = body
    return ids


@dataclass
class CaseInputs:
    """Decoder-level inputs of ``CaSETransformerSeqDecoder.forward`` (CaSE/Model.py:50), with
    ``source_map`` kept in its int64 index form (the one-hot of Utils.py:344-355 is never built)."""
    query: torch.Tensor            # int64 [B,1,Lq]
    passage: torch.Tensor          # int64 [B,NP,Lp]
    source_map: torch.Tensor       # int64 [B,S]   S = Lq + NP*Lp
    mem_q: torch.Tensor            # fp32  [B,1,Lq,H]
    mem_p: torch.Tensor            # fp32  [B,NP,Lp,H]
    prior_q: torch.Tensor          # fp32  [B,1,Lq]   (ones, Model.py:245)
    prior_p: torch.Tensor          # fp32  [B,NP,Lp]  (normalised, Model.py:239-243)
    answer_rep: torch.Tensor       # fp32  [B,H]      (Model.py:242)
    ids: torch.Tensor              # int64 [B]
    V: int = BERT_VOCAB
    meta: dict = field(default_factory=dict)

    @property
    def encode_memories(self) -> List[torch.Tensor]:
        return [self.mem_q, self.mem_p]

    @property
    def encode_masks(self) -> List[torch.Tensor]:
        return [self.query.ne(0), self.passage.ne(0)]

    @property
    def encode_weights(self) -> List[torch.Tensor]:
        return [self.prior_q, self.prior_p]

    def slice(self, lo: int, hi: int) -> "CaseInputs":
        f = lambda t: t[lo:hi]
        return CaseInputs(f(self.query), f(self.passage), f(self.source_map), f(self.mem_q), f(self.mem_p),
                          f(self.prior_q), f(self.prior_p), f(self.answer_rep), f(self.ids), self.V, dict(self.meta))

    def to(self, device, non_blocking=False) -> "CaseInputs":
        f = lambda t: t.to(device, non_blocking=non_blocking)
        return CaseInputs(f(self.query), f(self.passage), f(self.source_map), f(self.mem_q), f(self.mem_p),
                          f(self.prior_q), f(self.prior_p), f(self.answer_rep), f(self.ids), self.V, dict(self.meta))

    def pin(self) -> "CaseInputs":
        f = lambda t: t.pin_memory()
        return CaseInputs(f(self.query), f(self.passage), f(self.source_map), f(self.mem_q), f(self.mem_p),
                          f(self.prior_q), f(self.prior_p), f(self.answer_rep), f(self.ids), self.V, dict(self.meta))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.source_map, self.query, self.passage, self.mem_q, self.mem_p, self.prior_q,
                    self.prior_p, self.answer_rep))


def make_case_inputs(seed: int, B: int, Lq: int = 60, NP: int = 10, Lp: int = 100, V: int = BERT_VOCAB,
                     H: int = 256, id_base: int = 0, empty_frac: float = 0.1) -> CaseInputs:
    g = torch.Generator().manual_seed(seed)
    query = _token_rows(g, B, Lq, V, 0.5, 0.0, True).view(B, 1, Lq)
    passage = _token_rows(g, B * NP, Lp, V, 0.5, empty_frac, False).view(B, NP, Lp)
    source_map = torch.cat([query.reshape(B, -1), passage.reshape(B, -1)], dim=1)
    mem_q = torch.randn(B, 1, Lq, H, generator=g) * query.ne(0).unsqueeze(-1)
    mem_p = torch.randn(B, NP, Lp, H, generator=g) * passage.ne(0).unsqueeze(-1)
    answer_rep = torch.randn(B, H, generator=g)
    prior_p = torch.rand(B, NP, Lp, generator=g) * passage.ne(0)
    prior_p = prior_p / (1e-8 + prior_p.reshape(B, -1).sum(-1).view(B, 1, 1))
    prior_q = torch.ones(B, 1, Lq)
    ids = torch.arange(id_base, id_base + B)
    return CaseInputs(query, passage, source_map, mem_q, mem_p, prior_q, prior_p, answer_rep, ids, V,
                      dict(seed=seed, B=B, Lq=Lq, NP=NP, Lp=Lp, H=H))


@dataclass
class GttpInputs:
    """Step-side inputs of GTTP: what ``encode``/``init_decoder_states`` hand to ``decode``
    (GTTP/Model.py:156-181)."""
    context: torch.Tensor          # int64 [B,Lc]
    background: torch.Tensor       # int64 [B,Lb]
    background_map: torch.Tensor   # int64 [B,Lb]
    src_output: torch.Tensor       # fp32 [B,Lc,2H]   (c_enc_output)
    bg_output: torch.Tensor        # fp32 [B,Lb,2H]   (b_enc_output)
    init_state: torch.Tensor       # fp32 [B,1,H]
    ids: torch.Tensor
    V: int = 50000
    meta: dict = field(default_factory=dict)

    def slice(self, lo, hi):
        f = lambda t: t[lo:hi]
        return GttpInputs(f(self.context), f(self.background), f(self.background_map), f(self.src_output),
                          f(self.bg_output), f(self.init_state), f(self.ids), self.V, dict(self.meta))

    def to(self, device, non_blocking=False):
        f = lambda t: t.to(device, non_blocking=non_blocking)
        return GttpInputs(f(self.context), f(self.background), f(self.background_map), f(self.src_output),
                          f(self.bg_output), f(self.init_state), f(self.ids), self.V, dict(self.meta))


def make_gttp_inputs(seed: int, B: int, Lc: int = 60, NP: int = 10, Lp: int = 100, V: int = 50000,
                     H: int = 256, id_base: int = 0) -> GttpInputs:
    g = torch.Generator().manual_seed(seed)
    context = _token_rows(g, B, Lc, V, 0.5, 0.0, True)
    background = _token_rows(g, B * NP, Lp, V, 0.5, 0.1, False).view(B, NP * Lp)
    # gru_forward pads encoder outputs with zeros past each length (common/Utils.py:313-336);
    # pad_sequence-style layout keeps PADs only at the tail of each passage here.
    src_output = torch.randn(B, Lc, 2 * H, generator=g) * context.ne(0).unsqueeze(-1)
    bg_output = torch.randn(B, NP * Lp, 2 * H, generator=g) * background.ne(0).unsqueeze(-1)
    init_state = torch.randn(B, 1, H, generator=g) * 0.5
    ids = torch.arange(id_base, id_base + B)
    return GttpInputs(context, background, background.clone(), src_output, bg_output, init_state, ids, V,
                      dict(seed=seed, B=B, Lc=Lc, Lb=NP * Lp, H=H))


def make_vocab(V: int = BERT_VOCAB):
    """A synthetic vocabulary with the seven specials at their BERT ids (common/Constants.py:1-7);
    ``bert_tokenizer()`` (common/Utils.py:30-37) needs a download that is not available offline."""
    id2vocab = {i: f'tok{i}' for i in range(V)}
    for i, w in ((PAD, '[PAD]'), (BOS, '[unused0]'), (EOS, '[unused1]'), (UNK, '[UNK]'), (CLS, '[CLS]'),
                 (SEP, '[SEP]'), (103, '[MASK]')):
        if i < V:
            id2vocab[i] = w
    vocab2id = {w: i for i, w in id2vocab.items()}
    return vocab2id, id2vocab
