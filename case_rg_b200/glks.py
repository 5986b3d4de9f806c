"""GLKS on the same vocabulary-side kernels: ``Mixturer`` (GLKS/Model.py:135-147) mixes the vocabulary
distribution p_v [R,V] with the copy distribution p_k [R,Lb] scattered through the one-hot ``dyn_map`` [R,Lb,V] - the
a6 / a7 / a9 pattern of the CaSE step: softmax x gate, scatter-add on the int map, top-k.

``FastMixturer`` stands where GLKS builds ``Mixturer(hidden_size)`` (GLKS/Model.py:198): same constructor, same state_dict
keys (``linear1.weight`` [1,H], ``linear1.bias``), same ``forward(state, dists1, dists2, dyn_map)``; ``dyn_map`` may stay in
its int64 index form [R,Lb] (what GLKSDataset produces before ``build_map``, Utils.py:344-355) - the dense one-hot is
accepted and converted back.  The gate sigma(linear1(state)) is one H-long dot per row and stays a torch expression; the
[R,V] x gate scaling + scatter is ``case_copy_scatter`` (rows = hypotheses, W = 1, indices exact), the top-k of
``to_word`` is ``case_topk_rows`` (value desc, index asc = torch.topk / torch.max on ties).  ``FastVocabHead`` gives
``VocabGenerator``'s last two lines (generator Linear + softmax, GLKS/Model.py:128-130) on ``case_vocab_gemm`` +
``case_softmax_mix``.  ``install_fast_glks`` swaps both into a reference ``GLKS`` object.  No CPU fallback.
"""
import types

import torch
import torch.nn as nn

from . import _lib as L


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f'{what} needs CUDA tensors: there is no CPU fallback')


def topk_rows(dist: torch.Tensor, k: int, V: int = None):
    """Utils.topk (Utils.py:156-168): values [R,k] fp32, indices [R,k] int64; ties -> lower index first."""
    _require_cuda(dist, 'topk_rows')
    d = dist.float().contiguous()
    R, ld = d.shape
    V = ld if V is None else V
    if ld % 4:                                   # rows must be 16-byte aligned
        pad = torch.zeros(R, -(-ld // 4) * 4, device=d.device)
        pad[:, :ld] = d
        d, ld = pad, pad.size(1)
    vals = torch.empty(R, k, device=d.device)
    idx = torch.empty(R, k, dtype=torch.int32, device=d.device)
    with torch.cuda.device(d.device):
        L.call('case_topk_rows', d.data_ptr(), ld, R, V, k, vals.data_ptr(), idx.data_ptr(),
               torch.cuda.current_stream(d.device).cuda_stream)
    return vals, idx.long()


def copy_topk(gen_output, vocab_map, vocab_overlap, k=5):
    """Utils.copy_topk (Utils.py:170-178) on the device: fold the D dynamic entries of ``gen_output`` [R, V + D] onto
    their vocabulary ids, keep only the true out-of-vocabulary ones, top-k over the extended vocabulary.  ``vocab_map``:
    int [R, D] vocabulary id of every dynamic word (index form) or the reference's one-hot [R, D, V]; ``vocab_overlap``
    [R, D]."""
    _require_cuda(gen_output, 'copy_topk')
    if vocab_map.dim() == 3:
        V = vocab_map.size(-1)
        vocab_map = vocab_map.argmax(-1)
    else:
        V = gen_output.size(1) - vocab_map.size(1)
    R, D = vocab_map.shape
    ld = -(-(V + D) // 4) * 4
    g = torch.zeros(R, ld, device=gen_output.device)
    g[:, :V + D] = gen_output
    vm = vocab_map.to(torch.int32).contiguous()
    ov = vocab_overlap.float().contiguous()
    with torch.cuda.device(g.device):
        L.call('case_oov_fold', g.data_ptr(), ld, R, V, D, vm.data_ptr(), ov.data_ptr(), 1,
               torch.cuda.current_stream(g.device).cuda_stream)
    return topk_rows(g, k, V + D)


class FastMixturer(nn.Module):
    def __init__(self, hidden_size):
        super().__init__()
        self.linear1 = nn.Linear(hidden_size, 1)

    def forward(self, state, dists1, dists2, dyn_map):
        _require_cuda(dists1, 'FastMixturer')
        R, V = dists1.shape
        Lb = dists2.size(1)
        if dyn_map.dim() == 3:                    # one-hot from build_map: recover the indices
            dyn_map = dyn_map.argmax(dim=-1)
        p = torch.sigmoid(self.linear1(state.squeeze(1)))                    # [R,1]  (GLKS/Model.py:141)
        ld = -(-V // 4) * 4
        dist = torch.zeros(R, ld, device=dists1.device)
        dist[:, :V] = p * dists1
        fac = torch.zeros(R, 2, device=dists1.device)
        fac[:, :1] = 1.0 - p                      # F = 1 - p_k_v, M = 0: weight(r, s) = F * dists2[r, s] * exp(0 - 0)
        mp = dyn_map.to(torch.int32).contiguous()
        e = torch.zeros(R, Lb, device=dists1.device)
        e.masked_fill_(dists2 == 0, float('-inf'))                            # zero copy mass = masked source position
        pr = dists2.float().contiguous()
        with torch.cuda.device(dist.device):
            L.call('case_copy_scatter', mp.data_ptr(), Lb, 0, pr.data_ptr(), e.data_ptr(), fac.data_ptr(), 2,
                   dist.data_ptr(), ld, R, 1, Lb, V, torch.cuda.current_stream(dist.device).cuda_stream)
        return dist[:, :V]


class FastVocabHead(nn.Module):
    """``p = softmax(generator(feature))`` (GLKS/Model.py:128-130) on the vocabulary GEMM + softmax kernels."""

    def __init__(self, generator: nn.Linear):
        super().__init__()
        self.generator = generator

    def forward(self, feature):
        _require_cuda(feature, 'FastVocabHead')
        if feature.size(1) != L.H:
            raise ValueError(f'hidden size must be {L.H}')
        R, V = feature.size(0), self.generator.out_features
        ld = -(-V // 8) * 8
        logits = torch.empty(R, ld, device=feature.device)
        dist = torch.empty(R, ld, device=feature.device)
        gates = torch.ones(R, 4, device=feature.device)
        f = feature.float().contiguous()
        W = self.generator.weight.detach().float().contiguous()
        b = self.generator.bias.detach().float().contiguous()
        with torch.cuda.device(f.device):
            st = torch.cuda.current_stream(f.device).cuda_stream
            L.call('case_vocab_gemm', f.data_ptr(), W.data_ptr(), b.data_ptr(), logits.data_ptr(), R, V, ld, L.F32, 0, None, st)
            L.call('case_softmax_mix', logits.data_ptr(), ld, gates.data_ptr(), dist.data_ptr(), ld, R, V, 0, st)
        return dist[:, :V]


def install_fast_glks(model: nn.Module, device=None) -> nn.Module:
    """Swap the vocabulary side of a reference ``GLKS`` model (GLKS/Model.py:180-262) for the kernels, in place:
    ``model.mixture`` becomes a ``FastMixturer`` with the same weights, ``generate`` / ``to_word`` run on ``device`` with
    ``data['background_map']`` kept in index form (``forward`` no longer calls ``build_map``), ``VocabGenerator``'s
    generator + softmax go through ``FastVocabHead``.  Encoders, knowledge selection, attentions and the state tracker
    stay on the reference's own code (and device)."""
    mdev = next(model.parameters()).device
    dev = torch.device(device) if device is not None else mdev
    fast = FastMixturer(model.mixture.linear1.in_features)
    fast.load_state_dict(model.mixture.state_dict())
    model.mixture = fast.to(dev)
    vg = model.v_generator
    head = FastVocabHead(vg.generator).to(dev)
    orig_forward = model.forward

    def v_forward(self, p, word, state, segment, b_enc_output, c_enc_output, b_mask, c_mask):
        q = torch.cat([word, state, segment], dim=-1)
        c_output = self.c_attn(q, c_enc_output, c_enc_output, mask=c_mask.unsqueeze(1))[0].squeeze(1)
        b_output = self.b_attn(q, b_enc_output, b_enc_output, mask=b_mask.unsqueeze(1))[0].squeeze(1)
        feature = self.readout(torch.cat((word.squeeze(1), state.squeeze(1), segment.squeeze(1), c_output, b_output), dim=-1))
        return head(feature.to(dev))                                           # GLKS/Model.py:128-130 on the kernels

    def generate(self, data, encode_outputs, decode_outputs, softmax=True):
        p = self.mixture(decode_outputs['state'].to(dev), decode_outputs['p_v'].to(dev), decode_outputs['p_k'].to(dev),
                         data['background_map'].to(dev))
        return {'p': p}

    def to_word(self, data, gen_output, k=5, sampling=False):
        if sampling:
            raise NotImplementedError('sampling is not part of the test-mode path')
        vals, idx = topk_rows(gen_output['p'], k)
        return vals.to(mdev), idx.to(mdev)

    def forward(self, data, method='mle_train'):
        if method != 'test':
            return orig_forward(data, method=method)
        return {'answer': self.greedy(data) if self.beam_width == 1 else self.beam(data)}

    vg.forward = types.MethodType(v_forward, vg)
    model.generate = types.MethodType(generate, model)
    model.to_word = types.MethodType(to_word, model)
    model.forward = types.MethodType(forward, model)
    return model
