"""Oracle restatement of common/Generations.py (TEST INFRASTRUCTURE ONLY - see oracle/__init__.py).

A *stepper* stands in for the EncDecModel protocol (GTTP/EncDecModel.py:11-42): it owns the
encoder outputs and per-hypothesis decoder state and exposes

    stepper.B                               number of queries
    stepper.advance(parents, tokens) -> dist[n, V]

where row j of the call is a live hypothesis whose decoder input is ``tokens[j]`` and whose
state descends from row ``parents[j]`` of the previous call (``arange(B)`` on the first call).
That is exactly what ``decode`` + ``generate`` do to a fringe of ``Node``s
(Generations.py:147-158): ``get_data``/``concat_data`` (Utils.py:379-411) re-gather per-node state.
"""
import math
from typing import List

import torch

PAD, BOS, EOS, UNK = 0, 1, 2, 100


def topk(dist, k):
    """Utils.topk (Utils.py:156-168) without the never-used PAD/BOS/UNK zeroing."""
    if k > 1:
        return torch.topk(dist, k, dim=1, largest=True, sorted=True)
    return torch.max(dist, dim=1, keepdim=True)


def copy_topk(gen_output, vocab_map, vocab_overlap, k):
    """Utils.copy_topk (Utils.py:170-178): ``gen_output`` [R, V + D] = vocabulary entries then D dynamic entries;
    ``vocab_map`` one-hot [R, D, V] folds every dynamic entry onto its vocabulary id, ``vocab_overlap`` [R, D] keeps only
    the dynamic entries that are NOT vocabulary words; top-k over the V + D columns."""
    V = vocab_map.size(-1)
    vocab, dyn = gen_output[:, :V], gen_output[:, V:]
    vocab = vocab + torch.bmm(dyn.unsqueeze(1), vocab_map).squeeze(1)
    dyn = dyn * vocab_overlap
    return topk(torch.cat([vocab, dyn], dim=-1), k)


def greedy(stepper, max_len: int) -> torch.Tensor:
    """Generations.greedy (Generations.py:66-110): EOS at t==0 is rewritten to UNK but still ends
    the row; ended rows emit PAD (and PAD is what is fed back)."""
    B = stepper.B
    parents = torch.arange(B)
    inp = torch.full((B,), BOS, dtype=torch.long)
    ended = torch.zeros(B, dtype=torch.bool)
    out = []
    for t in range(max_len):
        dist = stepper.advance(parents, inp)
        _, ids = topk(dist, 1)
        tok = ids[:, 0].clone().cpu()        # (the stepper may evaluate on another device)
        this_end = tok == EOS
        if t == 0:
            tok[this_end] = UNK
        else:
            tok[ended] = PAD
        out.append(tok.unsqueeze(1))
        ended = ended | this_end
        inp = tok
    return torch.cat(out, dim=1)


class _Hyp:
    __slots__ = ('tokens', 'cum', 'length', 'q', 'row')

    def __init__(self, tokens, cum, length, q, row):
        self.tokens, self.cum, self.length, self.q, self.row = tokens, cum, length, q, row


def beam(stepper, max_len: int, width: int, return_all: bool = False):
    """Generations.beam (Generations.py:112-190).

    Semantics reproduced: costs are Python floats ``-log(p + 1e-10)`` accumulated from the root
    (Generations.py:170, Node.cum_cost :198); ranking key is cum_cost / length with length counting
    the BOS root (:199); per query the W best children over *all* its live parents are kept by a
    stable sort (:178-180); a kept child equal to EOS leaves the fringe at the next iteration and
    does not expand (:138-142), so the fringe can shrink below W; at l == max_len every surviving
    hypothesis is finished; the answer is the finished hypothesis with the smallest key, first
    finisher winning ties (:183-185); the returned row drops BOS and keeps EOS, zero-padded to
    the longest answer in the batch (:188, merge1D Utils.py:366-377).
    """
    B = stepper.B
    nxt: List[_Hyp] = [_Hyp([BOS], 0.0, 1, q, q) for q in range(B)]
    results = {q: [] for q in range(B)}
    for l in range(max_len + 1):
        fringe = []
        for h in nxt:
            if h.tokens[-1] == EOS or l == max_len:
                results[h.q].append(h)
            else:
                fringe.append(h)
        if not fringe:
            break
        parents = torch.tensor([h.row for h in fringe], dtype=torch.long)
        toks = torch.tensor([h.tokens[-1] for h in fringe], dtype=torch.long)
        dist = stepper.advance(parents, toks)
        probs, ids = topk(dist, width)
        probs, ids = probs.cpu(), ids.cpu()   # one transfer per step when the stepper evaluates on another device
        children = {q: [] for q in range(B)}
        for i, h in enumerate(fringe):
            for j in range(width):
                cost = -math.log(probs[i, j].item() + 1e-10)
                children[h.q].append(_Hyp(h.tokens + [ids[i, j].item()], h.cum + cost, h.length + 1, h.q, i))
        nxt = []
        for q in range(B):
            nxt += sorted(children[q], key=lambda n: n.cum / n.length)[:width]
    best = []
    for q in range(B):
        results[q].sort(key=lambda n: n.cum / n.length)
        best.append(results[q][0])
    L = max(len(h.tokens) - 1 for h in best)
    out = torch.zeros(B, L, dtype=torch.long)
    for q, h in enumerate(best):
        out[q, :len(h.tokens) - 1] = torch.tensor(h.tokens[1:], dtype=torch.long)
    if return_all:
        return out, best, results
    return out
