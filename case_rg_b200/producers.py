"""CaSE's pre-decode producers on the device: everything ``CaSE.do_test`` runs before the decoder
(CaSE/Model.py:313-331) - the shared ``TransformerSeqEncoder`` over query and passages
(common/TransformerSeqEncoderDecoder.py:14-45), ``Interaction`` (common/Interaction.py:15-76), the ``TransformerBlock``
stacks of passage selection and supporting-token identification (common/TransformerBlock.py:22-33, CaSE/Model.py:127-215)
and the prior / answer representation of ``ResponseGeneration.action`` (CaSE/Model.py:230-245).

``CaseProducers`` takes the full model's state_dict under the reference's key names ('query_encoder.*',
'passage_selection.*', 'span_extraction.*') and token ids; it returns what the decoder is handed: ``mem_q`` [B,1,Lq,H],
``mem_p`` [B,NP,Lp,H], ``prior_q``, ``prior_p``, ``answer_rep`` - plus ``rank`` (the passage scores ``do_test`` returns).
With it ``FastCaSE.search_ids`` decodes from token ids: the host ships a few hundred KB per batch instead of the 175 MB
of encoder outputs.

Kernels (csrc/producers.cu, csrc/gemm_rows.cu): embedding, LayerNorm (256 / 1280 wide), FlashAttention-style
self-attention with the key padding mask (head dims 32 and 160), the Interaction kernel (no [B*NP, Lp, Lq, 3H] tensor),
scorers, prior + answer representation, and ``case_gemm_rows_tc`` - a tcgen05 / TMEM GEMM with bias / activation /
residual / row-mask epilogues for every Linear.  bf16 GEMM operands, fp32 accumulation, statistics and residual
streams.  No CPU fallback.
"""
import math
import os
from typing import Dict

import torch

from . import _lib as L

H = L.H


def _u8(mask):
    return mask.to(torch.uint8).contiguous()


class _Linear:
    """nn.Linear weights for case_gemm_rows_tc: the [N, K] matrix packed per 64-wide K block into 256-row tiles of the
    UMMA canonical layout (engine.pack_vocab_tc), bias fp32; the plain bf16 matrix is kept for the cuBLAS A/B path."""

    def __init__(self, w, b, dev):
        from .engine import pack_vocab_tc
        w = w.detach().to(dev, torch.float32)
        self.N, self.K = w.shape
        if self.K % 64 or self.N % 256:
            raise ValueError(f'Linear needs in_features % 64 == 0 and out_features % 256 == 0, got {self.N} x {self.K}')
        self.w16 = w.to(torch.bfloat16).contiguous()
        self.wp = torch.stack([pack_vocab_tc(w[:, 64 * kb:64 * (kb + 1)], rows=256) for kb in range(self.K // 64)]).contiguous()
        self.b = (b.detach().to(dev, torch.float32) if b is not None else torch.zeros(self.N, device=dev)).contiguous()


class CaseProducers:
    def __init__(self, sd: Dict[str, torch.Tensor], device=None, prefix: str = '', gemm: str = None):
        if not torch.cuda.is_available():
            raise RuntimeError('case_rg_b200 needs a CUDA device: the producers have no CPU fallback')
        dev = torch.device(device if device is not None else 'cuda')
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        self.device = dev
        # 'tc' = the own tcgen05 GEMM (default); 'cublas' = torch.mm on the same bf16 operands (A/B and cross-check)
        self.gemm = gemm or os.environ.get('CASE_PRODUCER_GEMM', 'tc')
        g = lambda k: sd[prefix + k].detach().to(dev, torch.float32).contiguous()
        if g('query_encoder.embedding.0.weight').size(1) != H:
            raise ValueError(f'hidden_size must be {H}')
        self.E = g('query_encoder.embedding.0.weight')
        self.pe = g('query_encoder.embedding.1.pe')
        lin = lambda p: _Linear(sd[prefix + p + 'weight'], sd.get(prefix + p + 'bias'), dev)
        ln = lambda p: (g(p + 'weight'), g(p + 'bias'))

        def mha(p):
            return dict(inp=_Linear(sd[prefix + p + 'in_proj_weight'], sd[prefix + p + 'in_proj_bias'], dev),
                        out=lin(p + 'out_proj.'))
        self.enc = []
        nl = len({k[len(prefix):].split('.')[3] for k in sd if k.startswith(prefix + 'query_encoder.enc.layers.')})
        for l in range(nl):
            p = f'query_encoder.enc.layers.{l}.'
            self.enc.append(dict(att=mha(p + 'self_attn.'), l1=lin(p + 'linear1.'), l2=lin(p + 'linear2.'),
                                 n1=ln(p + 'norm1.'), n2=ln(p + 'norm2.')))

        def blocks(p):
            n = len({k[len(prefix + p):].split('.')[0] for k in sd if k.startswith(prefix + p)})
            return [dict(att=mha(f'{p}{i}.self_attn.'), l1=lin(f'{p}{i}.linear1.'), l2=lin(f'{p}{i}.linear2.'),
                         n1=ln(f'{p}{i}.norm1.'), n2=ln(f'{p}{i}.norm2.')) for i in range(n)]
        self.mods = {}
        for m in ('passage_selection.', 'span_extraction.'):
            self.mods[m] = dict(w=g(m + 'interaction.dual_att_linear.weight').reshape(-1).contiguous(),
                                qb=blocks(m + 'query_blocks.'), pb=blocks(m + 'passage_blocks.'),
                                sw=g(m + 'scorer.weight').reshape(-1).contiguous(), sb=g(m + 'scorer.bias').reshape(-1).contiguous())
        self.se_n1, self.se_n2 = ln('span_extraction.norm1.'), ln('span_extraction.norm2.')

    # ------------------------------------------------------------------ primitives (one launcher each)
    def _st(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _ln(self, x, p, C, add=None, want16=True, want32=False):
        M = x.size(0)
        y16 = torch.empty(M, C, dtype=torch.bfloat16, device=self.device) if want16 else None
        y32 = torch.empty(M, C, dtype=torch.float32, device=self.device) if want32 else None
        L.call('case_ln_rows_wide', x.data_ptr(), L.ptr(add), L.BF16 if x.dtype == torch.bfloat16 else L.F32, p[0].data_ptr(),
               p[1].data_ptr(), L.ptr(y16), L.ptr(y32), M, C, self._st())
        return y16, y32

    def _linear(self, x16, lin, act=0, residual=None, row_mask=None, out32=False):
        """y = act(x . W^T + b) (+ residual) (masked rows -> 0); act: 0 none, 1 gelu, 2 relu.  x16 bf16 [M, K]."""
        M = x16.size(0)
        if self.gemm == 'cublas':
            y = torch.mm(x16, lin.w16.t(), out_dtype=torch.float32) + lin.b
            if act == 1:
                y = torch.nn.functional.gelu(y)
            elif act == 2:
                y = torch.relu(y)
            if residual is not None:
                y = y + residual.float()
            if row_mask is not None:
                y = y * row_mask.view(-1, 1).float()
            return y if out32 else y.to(torch.bfloat16)
        y = torch.empty(M, lin.N, dtype=torch.float32 if out32 else torch.bfloat16, device=self.device)
        L.call('case_gemm_rows_tc', x16.data_ptr(), lin.wp.data_ptr(), lin.b.data_ptr(), M, lin.N, lin.K, act, L.ptr(residual),
               (L.BF16 if residual.dtype == torch.bfloat16 else L.F32) if residual is not None else 0, L.ptr(row_mask),
               y.data_ptr(), L.F32 if out32 else L.BF16, self._st())
        return y

    def _ffn(self, x16, l1, l2, act, residual=None, row_mask=None):
        """linear2(act(linear1(x))) (+ residual) (masked rows -> 0) -> fp32 [M, 256]: one launch (case_ffn_rows_tc), the hidden
        activations stay on the SM.  cuBLAS mode: two _linear calls."""
        if self.gemm == 'cublas' or l1.N != H or l2.N != H or l2.K != H:
            return self._linear(self._linear(x16, l1, act=act), l2, residual=residual, row_mask=row_mask, out32=True)
        M = x16.size(0)
        y = torch.empty(M, H, dtype=torch.float32, device=self.device)
        L.call('case_ffn_rows_tc', x16.data_ptr(), l1.wp.data_ptr(), l1.b.data_ptr(), l1.K, act, l2.wp.data_ptr(), l2.b.data_ptr(),
               M, L.ptr(residual), (L.BF16 if residual.dtype == torch.bfloat16 else L.F32) if residual is not None else 0,
               L.ptr(row_mask), y.data_ptr(), L.F32, self._st())
        return y

    def _attention(self, qkv, kmask, nseq, Lx, C):
        out = torch.empty(nseq * Lx, C, dtype=torch.bfloat16, device=self.device)
        L.call('case_enc_attention', qkv.data_ptr(), kmask.data_ptr(), nseq, Lx, C, 8, out.data_ptr(), self._st())
        return out

    # ------------------------------------------------------------------ modules
    def encode(self, ids):
        """TransformerSeqEncoder.forward: ids int [N, L] -> last-layer output fp32 [N * L, H]."""
        N, Lx = ids.shape
        M = N * Lx
        tok = ids.to(self.device).to(torch.int32).contiguous().view(-1)
        kmask = _u8(tok.ne(0))
        x = torch.empty(M, H, dtype=torch.float32, device=self.device)
        L.call('case_enc_embed', self.E.data_ptr(), self.pe.data_ptr(), tok.data_ptr(), M, Lx, math.sqrt(H), x.data_ptr(), self._st())
        for ly in self.enc:
            a16, a32 = self._ln(x, ly['n1'], H, want32=True)                       # src = norm1(src)
            att = self._attention(self._linear(a16, ly['att']['inp']), kmask, N, Lx, H)
            x = self._linear(att, ly['att']['out'], residual=a32, out32=True)      # src = src + attn
            b16, b32 = self._ln(x, ly['n2'], H, want32=True)                       # src = norm2(src)
            x = self._ffn(b16, ly['l1'], ly['l2'], 1, residual=b32)                # src = src + ffn
        return x

    def _block(self, bl, x, kmask, nseq, Lx):
        """TransformerBlock.forward: x bf16 [M, 1280] or fp32 [M, 256] -> fp32 [M, 256], PAD rows zero."""
        C = x.size(1)
        a16, _ = self._ln(x, bl['n1'], C)
        att = self._attention(self._linear(a16, bl['att']['inp']), kmask, nseq, Lx, C)
        r3 = self._linear(att, bl['att']['out'], residual=x, out32=(C == H))       # reps_temp3 = reps_temp1 + attention
        n16, _ = self._ln(r3, bl['n2'], C)
        return self._ffn(n16, bl['l1'], bl['l2'], 2, row_mask=kmask)

    def _interaction(self, w, Eq, Ep, qmask, pmask, B, NP, Lq, Lp):
        dev = self.device
        Gq_t = torch.empty(B * NP * Lq, 5 * H, dtype=torch.float32, device=dev)
        Gq = torch.empty(B * Lq, 5 * H, dtype=torch.bfloat16, device=dev)
        Gp = torch.empty(B * NP * Lp, 5 * H, dtype=torch.bfloat16, device=dev)
        L.call('case_interaction', Eq.data_ptr(), Ep.data_ptr(), qmask.data_ptr(), pmask.data_ptr(), w.data_ptr(), B, NP, Lq, Lp,
               Gq_t.data_ptr(), Gq.data_ptr(), Gp.data_ptr(), self._st())
        return Gq, Gp

    @torch.no_grad()
    def forward(self, query, passage):
        """query int [B, 1, Lq], passage int [B, NP, Lp] (any device) -> dict of device tensors."""
        with torch.cuda.device(self.device):
            return self._forward(query, passage)

    __call__ = forward

    def _forward(self, query, passage):
        dev = self.device
        B, NP, Lp = passage.shape
        Lq = query.size(2)
        query, passage = query.to(dev), passage.to(dev)
        qmask, pmask = _u8(query.reshape(-1).ne(0)), _u8(passage.reshape(-1).ne(0))
        enc_q = self.encode(query.reshape(B, Lq))
        enc_p = self.encode(passage.reshape(B * NP, Lp))
        out = {}
        q_rep, p_rep = enc_q, enc_p
        scores = {}
        for name in ('passage_selection.', 'span_extraction.'):
            m = self.mods[name]
            Gq, Gp = self._interaction(m['w'], q_rep, p_rep, qmask, pmask, B, NP, Lq, Lp)
            xq, xp = Gq, Gp
            for bl in m['qb']:
                xq = self._block(bl, xq, qmask, B, Lq)
            for bl in m['pb']:
                xp = self._block(bl, xp, pmask, B * NP, Lp)
            if name == 'passage_selection.':
                nrows, stride = B * NP, Lp                                         # scorer(passage_reps[:, :, 0]): the [CLS] rows
            else:
                nrows, stride = B * NP * Lp, 1                                     # scorer(passage_reps): every token
            sc = torch.empty(nrows, dtype=torch.float32, device=dev)
            L.call('case_rows_dot', xp.data_ptr(), m['sw'].data_ptr(), m['sb'].data_ptr(), nrows, stride, sc.data_ptr(), self._st())
            scores[name] = sc
            if name == 'passage_selection.':
                q_rep, p_rep = xq, xp
                ps_q, ps_p = xq, xp
            else:
                se_q, se_p = xq, xp
        _, mem_q = self._ln(ps_q, self.se_n1, H, add=se_q, want16=False, want32=True)   # norm1(query_rep + query_reps)
        _, mem_p = self._ln(ps_p, self.se_n2, H, add=se_p, want16=False, want32=True)
        prior_p = torch.empty(B, NP * Lp, dtype=torch.float32, device=dev)
        answer = torch.empty(B, H, dtype=torch.float32, device=dev)
        L.call('case_prior_answer', scores['passage_selection.'].data_ptr(), scores['span_extraction.'].data_ptr(), pmask.data_ptr(),
               mem_p.data_ptr(), B, NP, Lp, prior_p.data_ptr(), answer.data_ptr(), self._st())
        out.update(mem_q=mem_q.view(B, 1, Lq, H), mem_p=mem_p.view(B, NP, Lp, H), prior_q=torch.ones(B, 1, Lq, device=dev),
                   prior_p=prior_p.view(B, NP, Lp), answer_rep=answer, rank=scores['passage_selection.'].view(B, NP),
                   token_score=scores['span_extraction.'].view(B, NP, Lp), enc_q=enc_q.view(B, 1, Lq, H),
                   enc_p=enc_p.view(B, NP, Lp, H), ps_q=ps_q.view(B, 1, Lq, H), ps_p=ps_p.view(B, NP, Lp, H))
        return out
