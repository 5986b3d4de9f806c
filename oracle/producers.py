"""Oracle restatement of CaSE's pre-decode producers.  TEST INFRASTRUCTURE ONLY.

What ``CaSE.do_test`` runs before the decoder (CaSE/Model.py:313-331), function by function:

  TransformerSeqEncoder.forward ............... common/TransformerSeqEncoderDecoder.py:14-45
  TransformerEncoderLayer.forward ............. common/TransformerEncoder.py:54-77  (norm first, residual from the NORMALISED tensor)
  Interaction.forward ......................... common/Interaction.py:15-76
  TransformerBlock.forward .................... common/TransformerBlock.py:22-33
  RelevantPassageSelection.action ............. CaSE/Model.py:141-166
  SupportingTokenIdentification.action ........ CaSE/Model.py:188-215
  ResponseGeneration.action (prior, answer_rep) CaSE/Model.py:230-245

Plain torch fp32 over a state_dict with the reference's key names (the full ``CaSE`` model's: 'query_encoder.*',
'passage_selection.*', 'span_extraction.*').  ``nn.MultiheadAttention`` is evaluated the way torch evaluates it on
seq-first tensors with a key padding mask (``oracle.case_decoder._mha``).
"""
import math
from typing import Dict

import torch
import torch.nn.functional as F

from .case_decoder import _ln, _mha

NHEAD = 8


def encoder(sd, ids, prefix='query_encoder.'):
    """ids int64 [N, L] -> last-layer output [N, L, H] (PAD positions are NOT zeroed: the reference leaves them)."""
    H = sd[prefix + 'embedding.0.weight'].size(1)
    nl = len({k[len(prefix):].split('.')[2] for k in sd if k.startswith(prefix + 'enc.layers.')})
    mask = ids.ne(0)
    x = F.embedding(ids, sd[prefix + 'embedding.0.weight']) * math.sqrt(H) + sd[prefix + 'embedding.1.pe'][:ids.size(1)]
    x = x.transpose(0, 1)                                              # [L, N, H]
    for l in range(nl):
        p = f'{prefix}enc.layers.{l}.'
        x = _ln(x, sd, p + 'norm1')
        x = x + _mha(sd, p + 'self_attn', x, x, NHEAD, None, ~mask)
        x = _ln(x, sd, p + 'norm2')
        x = x + F.linear(F.gelu(F.linear(x, sd[p + 'linear1.weight'], sd[p + 'linear1.bias'])),
                         sd[p + 'linear2.weight'], sd[p + 'linear2.bias'])
    return x.transpose(0, 1)


def interaction(sd, prefix, E_q, E_p, q_mask, p_mask):
    """E_q [B, 1, Lq, H], E_p [B, NP, Lp, H], masks bool -> G_p_q [B, 1, Lq, 5H] (max over the passages), G_q_p
    [B, NP, Lp, 5H] (Interaction.py:15-76 with num_q = 1)."""
    B, NP, Lp, H = E_p.shape
    Lq = E_q.size(2)
    w = sd[prefix + 'dual_att_linear.weight'].view(3, H)
    Eq = E_q.expand(-1, NP, -1, -1).reshape(B * NP, Lq, H)
    Ep = E_p.reshape(B * NP, Lp, H)
    qm = q_mask.expand(-1, NP, -1).reshape(B * NP, Lq)
    pm = p_mask.reshape(B * NP, Lp)
    # U[i, j] = w . [E_q[j]; E_p[i]; E_q[j] * E_p[i]]
    U = (Eq @ w[0]).unsqueeze(1) + (Ep @ w[1]).unsqueeze(2) + torch.bmm(Ep * w[2], Eq.transpose(1, 2))
    mask = pm.unsqueeze(2) & qm.unsqueeze(1)
    U = U.masked_fill(~mask, float('-inf'))
    A = torch.softmax(U, dim=2).masked_fill(~mask, 0)
    Bm = torch.softmax(U, dim=1).masked_fill(~mask, 0)
    A1 = torch.bmm(A, Eq)                                             # [BN, Lp, H]
    B1 = torch.bmm(Bm.transpose(1, 2), Ep)                            # [BN, Lq, H]
    A2 = torch.bmm(A, B1)
    B2 = torch.bmm(Bm.transpose(1, 2), A1)
    G_q_p = torch.cat([Ep, A1, A2, Ep * A1, Ep * A2], -1).masked_fill(~pm.unsqueeze(-1), 0)
    G_p_q = torch.cat([Eq, B1, B2, Eq * B1, Eq * B2], -1).masked_fill(~qm.unsqueeze(-1), 0)
    G_q_p = G_q_p.view(B, NP, Lp, 5 * H)
    G_p_q = G_p_q.view(B, NP, Lq, 5 * H).max(dim=1, keepdim=True)[0]
    return G_p_q, G_q_p


def block(sd, prefix, x, mask):
    """TransformerBlock.forward (TransformerBlock.py:22-33): x [B, N, L, C] -> [B, N, L, H_out], masked rows zero."""
    B, N, L, C = x.shape
    r1 = x.reshape(-1, L, C)
    a = _ln(r1, sd, prefix + 'norm1').transpose(0, 1)
    att = _mha(sd, prefix + 'self_attn', a, a, NHEAD, None, ~mask.reshape(-1, L)).transpose(0, 1)
    r3 = r1 + att
    y = F.relu(F.linear(_ln(r3, sd, prefix + 'norm2'), sd[prefix + 'linear1.weight'], sd[prefix + 'linear1.bias']))
    y = F.linear(y, sd[prefix + 'linear2.weight'], sd[prefix + 'linear2.bias'])
    return y.view(B, N, L, -1).masked_fill(~mask.unsqueeze(-1), 0)


def _blocks(sd, prefix, x, mask):
    n = len({k[len(prefix):].split('.')[0] for k in sd if k.startswith(prefix)})
    for i in range(n):
        x = block(sd, f'{prefix}{i}.', x, mask)
    return x


def producers(sd: Dict[str, torch.Tensor], query, passage):
    """query int64 [B, 1, Lq], passage int64 [B, NP, Lp] -> everything ResponseGeneration.action hands the decoder
    (CaSE/Model.py:230-251) plus the intermediates the parity tests look at."""
    sd = {k: v.detach().float() for k, v in sd.items()}
    B, NP, Lp = passage.shape
    Lq = query.size(2)
    q_mask, p_mask = query.ne(0), passage.ne(0)
    enc_q = encoder(sd, query.reshape(-1, Lq)).view(B, 1, Lq, -1)
    enc_p = encoder(sd, passage.reshape(-1, Lp)).view(B, NP, Lp, -1)
    # ---- RelevantPassageSelection.action
    G_p_q, G_q_p = interaction(sd, 'passage_selection.interaction.', enc_q, enc_p, q_mask, p_mask)
    ps_q = _blocks(sd, 'passage_selection.query_blocks.', G_p_q, q_mask)
    ps_p = _blocks(sd, 'passage_selection.passage_blocks.', G_q_p, p_mask)
    passage_score = F.linear(ps_p[:, :, 0], sd['passage_selection.scorer.weight'], sd['passage_selection.scorer.bias']).squeeze(-1)
    # ---- SupportingTokenIdentification.action
    H_p_q, H_q_p = interaction(sd, 'span_extraction.interaction.', ps_q, ps_p, q_mask, p_mask)
    se_q = _blocks(sd, 'span_extraction.query_blocks.', H_p_q, q_mask)
    se_p = _blocks(sd, 'span_extraction.passage_blocks.', H_q_p, p_mask)
    token_score = F.linear(se_p, sd['span_extraction.scorer.weight'], sd['span_extraction.scorer.bias']).squeeze(-1)
    token_score = token_score.masked_fill(~p_mask, -1e6).clamp(min=-1e6, max=1e6)
    mem_q = _ln(ps_q + se_q, sd, 'span_extraction.norm1')
    mem_p = _ln(ps_p + se_p, sd, 'span_extraction.norm2')
    # ---- ResponseGeneration.action (the part before the decoder call)
    prior_p = torch.sigmoid(passage_score).unsqueeze(-1) * torch.sigmoid(token_score)
    prior_p = prior_p.reshape(B, -1)
    prior_p = prior_p / (1e-8 + prior_p.sum(dim=-1, keepdim=True))
    answer_rep = torch.bmm(prior_p.unsqueeze(1), mem_p.reshape(B, -1, mem_p.size(-1))).squeeze(1)
    prior_p = prior_p.view(B, NP, Lp)
    prior_q = torch.ones(B, 1, Lq, device=query.device)
    return dict(enc_q=enc_q, enc_p=enc_p, G_p_q=G_p_q, G_q_p=G_q_p, ps_q=ps_q, ps_p=ps_p, passage_score=passage_score,
                token_score=token_score, mem_q=mem_q, mem_p=mem_p, prior_q=prior_q, prior_p=prior_p, answer_rep=answer_rep)
