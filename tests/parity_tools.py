"""Search-level parity against the oracle at BASELINE sizes (test infrastructure; imports ``oracle``).

north_star's bar: greedy token sequences identical, beam results identical on >= 99 % of queries "allowing
tie-breaks".  A tie-break is made precise here by SCORING the CUDA path's answer with the oracle:

* greedy: rows are compared token by token; at the first position where a row differs, the oracle evaluates both
  candidates on the (identical) prefix.  The row is a *near-tie* iff the oracle's log-probability of the CUDA token is
  within ``tol_nats`` of the oracle's maximum - the two tokens are interchangeable at the storage precision.
* beam: the CUDA answer is teacher-forced through the oracle, giving its oracle key ``cum_cost / length``
  (Generations.py:198-199, costs -log(p + 1e-10)).  The query is a *near-tie* iff that key is within ``tol_nats``
  (nats per token) of the key of the oracle's own answer; if the CUDA answer's key is LOWER than the oracle's by
  more than that the search pruned differently and found a better hypothesis (reported as ``better``), if it is
  higher the query is a genuine ``miss``.

The tolerance is north_star's logit band turned into nats: the storage mode may move a logit by ``rel`` x max|logit|
(rel = 1e-4 for fp32 storage, 2e-2 for bf16 storage), so two alternatives whose oracle log-probabilities differ by
less than that are indistinguishable at that precision: ``tol_nats = rel * logit_scale`` with ``logit_scale`` the
oracle's own max|logit| on the case (``logit_scale`` below).  Every comparison also reports the largest gap it
saw, so the band actually used by the kernels is on record (profiles/).
"""
import math

import torch

from case_rg_b200 import synthetic as syn

PAD, BOS, EOS, UNK = syn.PAD, syn.BOS, syn.EOS, syn.UNK


def _pad(tokens, L):
    out = torch.zeros(tokens.size(0), L, dtype=torch.long)
    n = min(L, tokens.size(1))
    out[:, :n] = tokens[:, :n]
    return out


def logit_scale(stepper):
    """max |logit| (finite entries) of the oracle's last step - the scale north_star's relative logit band refers to."""
    lg = stepper.last['logits']
    return float(lg.masked_fill(~torch.isfinite(lg), 0).abs().max())


def oracle_greedy(stepper_factory, B, T):
    """Module-greedy tokens of the oracle (argmax, no EOS rule: CaSE/Model.py:119-122) plus, per step, the oracle's
    distribution rows (kept on the oracle's device) and the largest |logit| it saw."""
    st = stepper_factory()
    par, tok = torch.arange(B), torch.full((B,), BOS)
    toks, dists, scale = [], [], 0.0
    for _ in range(T):
        d = st.advance(par, tok)
        tok = d.argmax(1).cpu()
        toks.append(tok)
        dists.append(d)
        scale = max(scale, logit_scale(st))
    return torch.stack(toks, 1), dists, scale


def compare_greedy(got, want, dists, tol_nats):
    """-> dict(identical rows, near_tie rows, miss rows, decision counts, details).  ``got`` / ``want`` int64 [B,T] on the
    CPU; ``dists[t]`` the oracle's distribution at step t along ITS OWN greedy path (valid for a row up to and including
    the first step where the row differs).  The oracle's choice at a step is the argmax of its distribution (the proto-
    greedy rewrite of EOS to UNK at t = 0, Generations.py:99-100, changes the emitted token, not the choice)."""
    B, T = want.shape
    res = dict(rows=B, identical=0, near_tie=0, miss=0, decisions=0, flipped=0, max_gap_nats=0.0, tol_nats=tol_nats,
               first_diff=[], details=[])
    for b in range(B):
        neq = (got[b] != want[b]).nonzero()
        if neq.numel() == 0:
            res['identical'] += 1
            res['decisions'] += T
            continue
        t = int(neq[0])
        d = dists[t][b]
        p_top, p_got = float(d.max()), float(d[got[b, t]])
        gap = math.log(max(p_top, 1e-30) / max(p_got, 1e-30))
        res['decisions'] += t + 1          # comparable decisions of the row: the shared prefix plus the flipped one
        res['flipped'] += 1
        res['max_gap_nats'] = max(res['max_gap_nats'], gap)
        res['first_diff'].append(t)
        res['details'].append(dict(row=b, t=t, p_oracle_top=p_top, p_oracle_of_cuda_token=p_got, gap_nats=gap))
        if gap <= tol_nats:
            res['near_tie'] += 1
        else:
            res['miss'] += 1
    res['decision_agreement'] = 1.0 - res['flipped'] / max(1, res['decisions'])
    return res


def oracle_keys(stepper_factory, seqs, max_len):
    """Teacher-force every row of ``seqs`` (int64 [B,L], BOS dropped, EOS kept, zero padded) through the oracle ->
    (key = cum / length, cum, length) per row, with the reference's cost -log(p + 1e-10) taken from the oracle's fp32
    probability as a Python float (Generations.py:170)."""
    B, L = seqs.shape
    st = stepper_factory()
    par = torch.arange(B)
    tok = torch.full((B,), BOS)
    cum = [0.0] * B
    length = [1] * B
    done = [False] * B
    for t in range(min(L, max_len)):
        d = st.advance(par, tok)
        nxt = seqs[:, t]
        p = d[torch.arange(B, device=d.device), nxt.to(d.device)].cpu().tolist()
        for b in range(B):
            if done[b]:
                continue
            tk = int(nxt[b])
            if tk == PAD:                      # padding of a shorter answer (never a decoded token: EOS ends it first)
                done[b] = True
                continue
            cum[b] += -math.log(p[b] + 1e-10)
            length[b] += 1
            if tk == EOS:
                done[b] = True
        tok = nxt.clone()
    return [c / n for c, n in zip(cum, length)], cum, length


def compare_beam(stepper_factory, got, want, max_len, tol):
    """-> dict(queries, identical, near_tie, better, miss, details) for beam answers ``got`` (CUDA) / ``want`` (oracle);
    ``tol`` in nats per token."""
    B = want.size(0)
    L = max(got.size(1), want.size(1))
    got, want = _pad(got.cpu(), L), _pad(want.cpu(), L)
    same = (got == want).all(1)
    res = dict(queries=B, identical=int(same.sum()), near_tie=0, better=0, miss=0, max_gap_nats=0.0, tol_nats=tol,
               details=[])
    if bool(same.all()):
        return res
    k_got, _, _ = oracle_keys(stepper_factory, got, max_len)
    k_want, _, _ = oracle_keys(stepper_factory, want, max_len)
    for b in range(B):
        if bool(same[b]):
            continue
        gap = k_got[b] - k_want[b]
        res['max_gap_nats'] = max(res['max_gap_nats'], abs(gap))
        kind = 'near_tie' if abs(gap) <= tol else ('better' if gap < 0 else 'miss')
        res[kind] += 1
        res['details'].append(dict(query=b, key_cuda=k_got[b], key_oracle=k_want[b], gap=gap, kind=kind))
    return res


def finished_lengths(tokens):
    """Answer lengths (tokens up to and including EOS, or the row's non-zero length)."""
    out = []
    for row in tokens.tolist():
        n = 0
        for tk in row:
            if tk == PAD:
                break
            n += 1
            if tk == EOS:
                break
        out.append(n)
    return out
