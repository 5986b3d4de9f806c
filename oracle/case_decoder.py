"""Oracle restatement of the CaSE answer decoder (TEST INFRASTRUCTURE ONLY - see oracle/__init__.py).

Follows, function by function:
  CaSETransformerSeqDecoder.forward eval branch ........ CaSE/Model.py:50-63,91-125
  CaSETransformerSeqDecoder.extend ...................... CaSE/Model.py:38-48
  TransformerDecoderLayer.forward ....................... common/TransformerDecoder.py:61-90
  nn.MultiheadAttention (torch slow path: scale q, baddbmm(mask), softmax, bmm, out_proj)
                                                          called at TransformerDecoder.py:77,81
  BilinearAttention.matching/score/forward .............. common/BilinearAttention.py:13-60
  PositionalEmbedding.forward ........................... common/PositionalEmbedding.py:34-48
  generate_square_subsequent_mask ....................... common/Utils.py:23-28
  build_map (one-hot) ................................... common/Utils.py:344-355
  topk (k=1 -> torch.max) ............................... common/Utils.py:156-168

Everything is fp32 torch, on the CPU by default; weights come from a state_dict with the reference's key
names (prefix-free: 'embedding.0.weight', 'decs.0.layers.0.self_attn.in_proj_weight', ...).

``CaseOracle(sd, device=...)``: the full-size parity tests (BASELINE configs 2/4/5: B=64 x beam 4 x 40 steps over
2620 keys, V=30522) evaluate these same torch fp32 expressions on the GPU box's device so the checker finishes in
seconds; TF32 is switched off for that (``strict_fp32``), so the arithmetic stays IEEE fp32 like the CPU's.
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

NEG_CAUSAL = -1e20   # neginf(float32), common/Utils.py:14-21


def strict_fp32():
    """fp32 matmuls stay fp32 when the oracle is evaluated on a CUDA device (no TF32)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision('highest')


def _ln(x, sd, name):
    return F.layer_norm(x, (x.size(-1),), sd[name + '.weight'], sd[name + '.bias'], 1e-5)


def _mha(sd, name, query, memory, nhead, attn_mask=None, key_padding_mask=None):
    """nn.MultiheadAttention.forward on seq-first tensors, the way torch evaluates it when
    need_weights=True (the reference always takes the weights, TransformerDecoder.py:77,81).
    query [Tq,R,H]; memory [Tk,R,H]; attn_mask float [Tq,Tk]; key_padding_mask bool [R,Tk] True=pad."""
    Tq, R, H = query.shape
    Tk = memory.size(0)
    hd = H // nhead
    W, b = sd[name + '.in_proj_weight'], sd[name + '.in_proj_bias']
    q = F.linear(query, W[:H], b[:H])
    k = F.linear(memory, W[H:2 * H], b[H:2 * H])
    v = F.linear(memory, W[2 * H:], b[2 * H:])
    q = q.reshape(Tq, R * nhead, hd).transpose(0, 1) * math.sqrt(1.0 / hd)
    k = k.reshape(Tk, R * nhead, hd).transpose(0, 1)
    v = v.reshape(Tk, R * nhead, hd).transpose(0, 1)
    mask = None
    if key_padding_mask is not None:
        kp = torch.zeros(R, Tk, device=query.device).masked_fill(key_padding_mask, float('-inf'))
        mask = kp.view(R, 1, 1, Tk).expand(-1, nhead, -1, -1).reshape(R * nhead, 1, Tk)
    if attn_mask is not None:
        mask = attn_mask if mask is None else attn_mask + mask
    scores = torch.bmm(q, k.transpose(1, 2)) if mask is None else torch.baddbmm(mask, q, k.transpose(1, 2))
    w = torch.softmax(scores, dim=-1)
    o = torch.bmm(w, v).transpose(0, 1).reshape(Tq * R, H)
    o = F.linear(o, sd[name + '.out_proj.weight'], sd[name + '.out_proj.bias'])
    return o.view(Tq, R, H)


def _decoder_layer(sd, p, x, memory, nhead, tgt_mask, tgt_kpm, mem_kpm):
    """TransformerDecoderLayer.forward (TransformerDecoder.py:76-89): note the residual is taken
    from the *normalised* tensor (norm first, then x + sublayer(x))."""
    x = _ln(x, sd, p + 'norm1')
    x = x + _mha(sd, p + 'self_attn', x, x, nhead, tgt_mask, tgt_kpm)
    x = _ln(x, sd, p + 'norm2')
    x = x + _mha(sd, p + 'multihead_attn', x, memory, nhead, None, mem_kpm)
    x = _ln(x, sd, p + 'norm3')
    y = F.linear(F.gelu(F.linear(x, sd[p + 'linear1.weight'], sd[p + 'linear1.bias'])),
                 sd[p + 'linear2.weight'], sd[p + 'linear2.bias'])
    return x + y


def _additive_attention(sd, name, query, key, value, mask):
    """BilinearAttention.forward (BilinearAttention.py:24-60): v . tanh(Wq q + b + Uk k),
    masked_fill(-inf), softmax, masked_fill(0), context = attn @ value.
    query [R,t,Dq]; key/value [R,S,Dk]; mask bool [R,t,S]."""
    wq = F.linear(query, sd[name + '.linear_query.weight'], sd[name + '.linear_query.bias']).unsqueeze(-2)
    uh = F.linear(key, sd[name + '.linear_key.weight']).unsqueeze(-3)
    e = F.linear(torch.tanh(wq + uh), sd[name + '.v.weight']).squeeze(-1)
    e = e.masked_fill(~mask, float('-inf'))
    a = torch.softmax(e, dim=-1).masked_fill(~mask, 0)
    return torch.bmm(a, value), a


def causal_mask(n, device=None):
    m = torch.tril(torch.ones(n, n, dtype=torch.bool, device=device))
    return torch.zeros(n, n, device=device).masked_fill(~m, NEG_CAUSAL)


def onehot_map(source_map, V):
    """build_map (Utils.py:344-355): dense fp32 one-hot [B,S,V]."""
    B, S = source_map.shape
    m = torch.zeros(B, S, V, device=source_map.device)
    m.scatter_(2, source_map.unsqueeze(2), 1.0)
    return m


class CaseOracle:
    """Reference-order evaluation of the decoder over a whole token prefix."""

    def __init__(self, sd: Dict[str, torch.Tensor], nhead: int = 8, device=None):
        self.dev = torch.device(device if device is not None else 'cpu')
        if self.dev.type == 'cuda':
            strict_fp32()
        self.sd = {k: v.detach().float().to(self.dev) for k, v in sd.items()}
        self.H = sd['embedding.0.weight'].size(1)
        self.V = sd['gen.2.weight'].size(0)
        self.nhead = nhead
        self.M = len({k.split('.')[1] for k in sd if k.startswith('decs.')})
        self.L = len({k.split('.')[3] for k in sd if k.startswith('decs.0.layers.')})

    # ---- Model.py:56-58: flatten memories / masks / weights to [B, S_i(, H)]
    def prepare(self, inp):
        B = inp.source_map.size(0)
        d = self.dev
        return dict(
            B=B,
            mem=[m.reshape(B, -1, self.H).float().to(d) for m in inp.encode_memories],
            mask=[m.reshape(B, -1).to(d) for m in inp.encode_masks],
            w=[w.reshape(B, -1).float().to(d) for w in inp.encode_weights],
            feat=inp.answer_rep.float().to(d),
            source_map=inp.source_map.to(d),
        )

    def embed(self, idx):
        """embedding + PositionalEmbedding (Model.py:96; PositionalEmbedding.py:44-48)."""
        x = F.embedding(idx, self.sd['embedding.0.weight'])
        return x * math.sqrt(self.H) + self.sd['embedding.1.pe'][:idx.size(1)].unsqueeze(0)

    def prefix_forward(self, ctx, idx, row2q=None, onehot: Optional[torch.Tensor] = None):
        """Body of the eval loop for one prefix (Model.py:95-117).  idx int64 [R,n] (BOS first).
        Returns every intermediate the parity tests look at, for all n positions."""
        sd, H = self.sd, self.H
        idx = idx.to(self.dev)
        R, n = idx.shape
        if row2q is None:
            row2q = torch.arange(R, device=self.dev)
        x_in = self.embed(idx)                                              # [R,n,H]
        feat = _ln(ctx['feat'][row2q], sd, 'norm2').unsqueeze(1).expand(-1, n, -1)
        h = x_in.transpose(0, 1)                                            # [n,R,H]
        tok_valid = idx.ne(0)
        cm, ps, attn_raw = [], [], []
        for i in range(self.M):
            mem = ctx['mem'][i][row2q]
            mmask = ctx['mask'][i][row2q]
            for l in range(self.L):
                h = _decoder_layer(sd, f'decs.{i}.layers.{l}.', h, mem.transpose(0, 1), self.nhead,
                                   causal_mask(n, self.dev), ~tok_valid, ~mmask)
            # Model.py:108: mask = outer(idx != 0, mem_mask)
            m2 = tok_valid.unsqueeze(-1) & mmask.unsqueeze(1)
            c, a = _additive_attention(sd, f'attns.{i}', torch.cat([h.transpose(0, 1), feat], -1), mem, mem, m2)
            p = ctx['w'][i][row2q].unsqueeze(1) * a                         # Model.py:110
            p = p / (1e-8 + p.sum(-1, keepdim=True))                        # Model.py:111
            cm.append(c); ps.append(p); attn_raw.append(a)
        hN = _ln(h, sd, 'norm1').transpose(0, 1)                            # Model.py:113
        f = F.linear(torch.cat([x_in, hN, feat], -1), sd['gen.0.weight'], sd['gen.0.bias'])
        logits = F.linear(f, sd['gen.2.weight'])
        gen = torch.softmax(logits, -1)                                     # Model.py:115
        gates = torch.softmax(F.linear(torch.cat([hN] + cm, -1), sd['mix.weight'], sd['mix.bias']), -1)
        copy_w = torch.cat([gates[:, :, i + 1].unsqueeze(-1) * ps[i] for i in range(self.M)], -1)  # Model.py:42
        smap = ctx['source_map'][row2q]
        if onehot is not None:
            copy = torch.bmm(copy_w, onehot[row2q])                         # Model.py:43
        else:
            copy = torch.zeros(R, n, self.V, device=self.dev).scatter_add_(2, smap.unsqueeze(1).expand(-1, n, -1), copy_w)
        dist = gates[:, :, 0].unsqueeze(-1) * gen + copy                    # Model.py:41,48
        return dict(dist=dist, gen=gen, logits=logits, gates=gates, p=ps, attn=attn_raw, ctx=cm,
                    dec_out=hN, gen_feat=f, copy_w=copy_w)

    def greedy_module(self, inp, T: int, dense_onehot: bool = True):
        """The in-module greedy loop (Model.py:91-123): no EOS test, whole prefix recomputed each
        step, argmax via torch.max (first index on ties).  Returns tokens [B,T] and the last
        step's outputs (which cover all T positions, as the reference returns them)."""
        ctx = self.prepare(inp)
        B = ctx['B']
        oh = onehot_map(ctx['source_map'], self.V) if dense_onehot else None
        idx = torch.full((B, 1), 1, dtype=torch.long, device=self.dev)
        outs = []
        last = None
        for _ in range(T):
            last = self.prefix_forward(ctx, idx, onehot=oh)
            _, y = torch.max(last['dist'][:, -1], dim=1, keepdim=True)
            outs.append(y)
            idx = torch.cat([idx, y], dim=1)
        return torch.cat(outs, dim=1), last

    # ---- protocol face used by oracle.generations (the adapter of SURVEY.md §8c, restated)
    def stepper(self, inp, dense_onehot: bool = False):
        return _PrefixStepper(self, inp, dense_onehot)

    def incremental(self, inp, n_oov: int = 0):
        return IncrementalStepper(self, inp, n_oov)


class _PrefixStepper:
    """Hypothesis state = its token prefix; every advance recomputes the prefix (reference cost)."""

    def __init__(self, orc: CaseOracle, inp, dense_onehot):
        self.o, self.ctx = orc, orc.prepare(inp)
        self.oh = onehot_map(self.ctx['source_map'], orc.V) if dense_onehot else None
        self.B = self.ctx['B']
        self.prefix = torch.zeros(self.B, 0, dtype=torch.long, device=orc.dev)
        self.row2q = torch.arange(self.B, device=orc.dev)
        self.last = None

    def advance(self, parents, tokens):
        parents, tokens = parents.to(self.o.dev), tokens.to(self.o.dev)
        self.prefix = torch.cat([self.prefix[parents], tokens.view(-1, 1)], dim=1)
        self.row2q = self.row2q[parents]
        self.last = self.o.prefix_forward(self.ctx, self.prefix, self.row2q, self.oh)
        return self.last['dist'][:, -1]


class IncrementalStepper:
    """KV-cached evaluation of the same function: by causality (SURVEY.md §3.2 probe) position j's
    hidden state never changes once computed, so only the newest position is evaluated and the
    per-layer self-attention K/V rows are cached.  Cross-attention K/V and Uk.mem are projected
    once.  This is the executable spec of the CUDA path's data flow; it is validated against
    ``prefix_forward`` in tests/test_oracle_golden.py."""

    def __init__(self, orc: CaseOracle, inp, n_oov: int = 0):
        """n_oov > 0: pointer-generator OOV extension of CaSE/Model.py:38-48 - ``source_map`` ids in [V, V + n_oov) are
        per-query dynamic words: the distribution has V + n_oov columns (the one-hot of build_map simply gets that many,
        Utils.py:344-355 with max = V + n_oov; the generation part is zero there) and a dynamic id is fed back as UNK,
        having no embedding row."""
        self.o = orc
        self.n_oov = int(n_oov)
        sd, H = orc.sd, orc.H
        self.ctx = ctx = orc.prepare(inp)
        self.B = ctx['B']
        self.feat = _ln(ctx['feat'], sd, 'norm2')
        self.xk, self.xv, self.uk = {}, {}, []
        for i in range(orc.M):
            for l in range(orc.L):
                W = sd[f'decs.{i}.layers.{l}.multihead_attn.in_proj_weight']
                b = sd[f'decs.{i}.layers.{l}.multihead_attn.in_proj_bias']
                self.xk[i, l] = F.linear(ctx['mem'][i], W[H:2 * H], b[H:2 * H])
                self.xv[i, l] = F.linear(ctx['mem'][i], W[2 * H:], b[2 * H:])
            self.uk.append(F.linear(ctx['mem'][i], sd[f'attns.{i}.linear_key.weight']))
        dev = orc.dev
        self.row2q = torch.arange(self.B, device=dev)
        self.t = 0
        self.tokens = torch.zeros(self.B, 0, dtype=torch.long, device=dev)
        self.kc = {(i, l): torch.zeros(self.B, 0, H, device=dev) for i in range(orc.M) for l in range(orc.L)}
        self.vc = {(i, l): torch.zeros(self.B, 0, H, device=dev) for i in range(orc.M) for l in range(orc.L)}
        self.last = None

    def advance(self, parents, tokens):
        o, sd, H, nh = self.o, self.o.sd, self.o.H, self.o.nhead
        hd = H // nh
        parents, tokens = parents.to(o.dev), tokens.to(o.dev)
        self.row2q = self.row2q[parents]
        self.tokens = torch.cat([self.tokens[parents], tokens.view(-1, 1)], dim=1)
        for k in self.kc:
            self.kc[k] = self.kc[k][parents]
            self.vc[k] = self.vc[k][parents]
        R, t = tokens.numel(), self.t
        q2 = self.row2q
        tokens = torch.where(tokens >= o.V, torch.full_like(tokens, 100), tokens) if self.n_oov else tokens   # OOV -> UNK
        x_in = F.embedding(tokens, sd['embedding.0.weight']) * math.sqrt(H) + sd['embedding.1.pe'][t]
        feat = self.feat[q2]
        valid = tokens.ne(0)
        key_pad = self.tokens.eq(0)                                           # [R,t+1]
        h = x_in
        cm, ps = [], []
        scale = math.sqrt(1.0 / hd)
        for i in range(o.M):
            mmask = self.ctx['mask'][i][q2]
            for l in range(o.L):
                p = f'decs.{i}.layers.{l}.'
                a = _ln(h, sd, p + 'norm1')
                W, b = sd[p + 'self_attn.in_proj_weight'], sd[p + 'self_attn.in_proj_bias']
                qkv = F.linear(a, W, b)
                self.kc[i, l] = torch.cat([self.kc[i, l], qkv[:, None, H:2 * H]], 1)
                self.vc[i, l] = torch.cat([self.vc[i, l], qkv[:, None, 2 * H:]], 1)
                q = (qkv[:, :H] * scale).view(R, nh, 1, hd)
                K = self.kc[i, l].view(R, t + 1, nh, hd).transpose(1, 2)
                Vv = self.vc[i, l].view(R, t + 1, nh, hd).transpose(1, 2)
                s = (q @ K.transpose(-1, -2)).masked_fill(key_pad[:, None, None, :], float('-inf'))
                c = (torch.softmax(s, -1) @ Vv).reshape(R, H)
                h = a + F.linear(c, sd[p + 'self_attn.out_proj.weight'], sd[p + 'self_attn.out_proj.bias'])
                bb = _ln(h, sd, p + 'norm2')
                W, b = sd[p + 'multihead_attn.in_proj_weight'], sd[p + 'multihead_attn.in_proj_bias']
                q = (F.linear(bb, W[:H], b[:H]) * scale).view(R, nh, 1, hd)
                K = self.xk[i, l][q2].view(R, -1, nh, hd).transpose(1, 2)
                Vv = self.xv[i, l][q2].view(R, -1, nh, hd).transpose(1, 2)
                s = (q @ K.transpose(-1, -2)).masked_fill(~mmask[:, None, None, :], float('-inf'))
                c = (torch.softmax(s, -1) @ Vv).reshape(R, H)
                h = bb + F.linear(c, sd[p + 'multihead_attn.out_proj.weight'], sd[p + 'multihead_attn.out_proj.bias'])
                cc = _ln(h, sd, p + 'norm3')
                h = cc + F.linear(F.gelu(F.linear(cc, sd[p + 'linear1.weight'], sd[p + 'linear1.bias'])),
                                  sd[p + 'linear2.weight'], sd[p + 'linear2.bias'])
            qa = F.linear(torch.cat([h, feat], -1), sd[f'attns.{i}.linear_query.weight'],
                          sd[f'attns.{i}.linear_query.bias'])
            e = (torch.tanh(qa.unsqueeze(1) + self.uk[i][q2]) @ sd[f'attns.{i}.v.weight'].view(H, 1)).squeeze(-1)
            m2 = valid.unsqueeze(-1) & mmask
            e = e.masked_fill(~m2, float('-inf'))
            a = torch.softmax(e, -1).masked_fill(~m2, 0)
            cm.append(torch.bmm(a.unsqueeze(1), self.ctx['mem'][i][q2]).squeeze(1))
            pw = self.ctx['w'][i][q2] * a
            ps.append(pw / (1e-8 + pw.sum(-1, keepdim=True)))
        hN = _ln(h, sd, 'norm1')
        f = F.linear(torch.cat([x_in, hN, feat], -1), sd['gen.0.weight'], sd['gen.0.bias'])
        logits = F.linear(f, sd['gen.2.weight'])
        gen = torch.softmax(logits, -1)
        gates = torch.softmax(F.linear(torch.cat([hN] + cm, -1), sd['mix.weight'], sd['mix.bias']), -1)
        copy_w = torch.cat([gates[:, i + 1:i + 2] * ps[i] for i in range(o.M)], -1)
        dist = gates[:, :1] * gen
        if self.n_oov:
            dist = torch.cat([dist, dist.new_zeros(R, self.n_oov)], 1)
        dist = dist.scatter_add(1, self.ctx['source_map'][q2], copy_w)
        self.t += 1
        self.last = dict(dist=dist, gen=gen, logits=logits, gates=gates, p=ps, ctx=cm, dec_out=hN,
                         gen_feat=f, copy_w=copy_w, x_in=x_in)
        return dist
