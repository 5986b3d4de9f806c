"""Oracle restatement of the GTTP pointer-generator decode step (TEST INFRASTRUCTURE ONLY).

Follows:
  BBCDecoder.forward ................ GTTP/Model.py:113-131
  CopyGenerator.forward ............. GTTP/Model.py:14-43
  GTTP.decode / generate / to_word .. GTTP/Model.py:176-193
  BilinearAttention ................. common/BilinearAttention.py:13-60
  nn.GRU single step, gate order r,z,n (torch): r=s(Wir x+bir+Whr h+bhr), z=s(...),
      n=tanh(Win x+bin + r*(Whn h+bhn)), h'=(1-z)*n + z*h

The encoders (bi-GRUs, GTTP/Model.py:156-174) are outside the hot path: their outputs are inputs.
"""
from typing import Dict

import torch
import torch.nn.functional as F

from .case_decoder import _additive_attention, onehot_map


class GttpOracle:
    def __init__(self, sd: Dict[str, torch.Tensor], device=None):
        self.dev = torch.device(device if device is not None else 'cpu')
        if self.dev.type == 'cuda':
            from .case_decoder import strict_fp32
            strict_fp32()
        self.sd = {k: v.detach().float().to(self.dev) for k, v in sd.items()}
        self.H = sd['dec.gru.weight_hh_l0'].size(1)
        self.V = sd['gen.linear.weight'].size(0)

    def step(self, tok, state, src_output, bg_output, c_mask, b_mask):
        """decode (Model.py:176-180) -> feature [R,H], new state [R,1,H], bg_attn [R,Lb]."""
        sd, H = self.sd, self.H
        emb = F.embedding(tok, sd['dec.embedding.weight'])
        q = state[:, -1].unsqueeze(1)
        src_ctx, src_a = _additive_attention(sd, 'dec.src_attn', q, src_output, src_output, c_mask.unsqueeze(1))
        bg_ctx, bg_a = _additive_attention(sd, 'dec.bg_attn', q, bg_output, bg_output, b_mask.unsqueeze(1))
        src_ctx, bg_ctx, bg_a = src_ctx.squeeze(1), bg_ctx.squeeze(1), bg_a.squeeze(1)
        x = torch.cat([emb, src_ctx, bg_ctx], 1)
        h = state[:, -1]
        gi = F.linear(x, sd['dec.gru.weight_ih_l0'], sd['dec.gru.bias_ih_l0'])
        gh = F.linear(h, sd['dec.gru.weight_hh_l0'], sd['dec.gru.bias_hh_l0'])
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        hn = (1 - z) * n + z * h
        feat = F.linear(torch.cat([emb, hn, src_ctx, bg_ctx], 1), sd['dec.readout.weight'], sd['dec.readout.bias'])
        return feat, hn.unsqueeze(1), bg_a

    def generate(self, feat, bg_attn, bmap, onehot=None):
        """CopyGenerator.forward (Model.py:25-43); ``bmap`` int64 [R,Lb]."""
        sd = self.sd
        logits = F.linear(feat, sd['gen.linear.weight'], sd['gen.linear.bias'])
        logits[:, 0] = float('-inf')
        gen = torch.softmax(logits, 1)
        p_copy = torch.sigmoid(F.linear(feat, sd['gen.linear_copy.weight'], sd['gen.linear_copy.bias']))
        if onehot is not None:
            copy = torch.bmm(bg_attn.unsqueeze(1), onehot).squeeze(1)
        else:
            copy = torch.zeros_like(gen).scatter_add_(1, bmap, bg_attn)
        return gen * (1 - p_copy) + copy * p_copy, dict(gen=gen, p_copy=p_copy, logits=logits)

    def stepper(self, inp, dense_onehot=False):
        return _GttpStepper(self, inp, dense_onehot)


class _GttpStepper:
    def __init__(self, orc, inp, dense_onehot):
        self.o, self.inp = orc, inp.to(orc.dev)
        inp = self.inp
        self.B = inp.context.size(0)
        self.state = inp.init_state.float()
        self.row2q = torch.arange(self.B, device=orc.dev)
        self.oh = onehot_map(inp.background_map, orc.V) if dense_onehot else None
        self.last = None

    def advance(self, parents, tokens):
        inp = self.inp
        parents, tokens = parents.to(self.o.dev), tokens.to(self.o.dev)
        self.row2q = self.row2q[parents]
        q2 = self.row2q
        feat, st, bg_a = self.o.step(tokens, self.state[parents], inp.src_output[q2], inp.bg_output[q2],
                                     inp.context[q2].ne(0), inp.background[q2].ne(0))
        self.state = st
        dist, aux = self.o.generate(feat, bg_a, inp.background_map[q2], None if self.oh is None else self.oh[q2])
        self.last = dict(dist=dist, feat=feat, bg_attn=bg_a, state=st, **aux)
        return dist
