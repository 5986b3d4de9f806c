#!/usr/bin/env python
"""Golden vectors for case_rg_b200/results.py, produced by the UNMODIFIED reference functions
(common/Utils.py: to_sentence, remove_duplicate, bert_detokenizer; Utils.py: save_result).
Runs only in the build container (imports /root/reference); writes results_golden.json next to itself.

    python tests/golden/make_results_golden.py
"""
import json
import math
import os
import random
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def install_reference():
    for m in ('bcolz', 'nltk'):
        sys.modules[m] = types.ModuleType(m)
    tr = types.ModuleType('transformers')
    tr.torch, tr.math = torch, math
    tr.__all__ = ['torch', 'math']
    sys.modules['transformers'] = tr
    sys.path.insert(0, '/root/reference')


install_reference()
import common.Utils as CU      # noqa: E402
import Utils as TU             # noqa: E402  (reference top-level Utils.py: save_result)


def vocab(V=64):
    words = ['[PAD]', '[unused0]', '[unused1]'] + ['w%d' % i for i in range(3, V)]
    words[5], words[6], words[7] = '##ing', '[UNK]', '##s'
    return {i: w for i, w in enumerate(words)}


class Dataset:
    """the four accessors save_result uses (CaSE/CaSEDataset.py:118-128)"""

    def __init__(self, n, pool):
        self.samples = [dict(context_id=['c%d_%d' % (i, j) for j in range(i % 3)], query_id='q%d' % i,
                             passage_id=['p%d_%d' % (i, j) for j in range(1 + i % 2)],
                             passage_pool_id=['d%d_%d' % (i, j) for j in range(pool)]) for i in range(n)]

    def context_id(self, id): return self.samples[id]['context_id']
    def query_id(self, id): return self.samples[id]['query_id']
    def passage_id(self, id): return self.samples[id]['passage_id']
    def pool(self, id): return self.samples[id]['passage_pool_id']


def main():
    rnd = random.Random(7)
    id2vocab = vocab()
    V = len(id2vocab)
    # token rows: short vocab ranges so repeats (and therefore cuts) are frequent, BOS/PAD/EOS sprinkled in
    rows = []
    for k in range(200):
        L = rnd.choice([1, 2, 3, 4, 6, 9, 14, 20, 40])
        hi = rnd.choice([4, 6, 9, 16, V])
        r = [rnd.randrange(0, hi) for _ in range(L)]
        if k % 5 == 0:
            r[rnd.randrange(L)] = 2
        rows.append(r + [0] * (40 - L))
    tokens = torch.tensor(rows, dtype=torch.int64)
    sents = CU.to_sentence(tokens, id2vocab)
    words_before = [list(s) for s in sents]
    CU.remove_duplicate(sents)
    detok = CU.bert_detokenizer()
    answers = [detok(s) for s in sents]
    # remove_duplicate on hand-made lists with other n
    extra = []
    for n in (1, 2, 3, 5):
        for _ in range(60):
            L = rnd.randrange(0, 18)
            s = [rnd.randrange(0, rnd.choice([2, 3, 5, 8])) for _ in range(L)]
            t = [list(s)]
            CU.remove_duplicate(t, n)
            extra.append(dict(n=n, sent=s, out=t[0]))
    # save_result: two batches, answers + rank scores with ties
    ds = Dataset(12, pool=5)
    preds = []
    g = torch.Generator().manual_seed(3)
    for b0 in (0, 6):
        ids = torch.arange(b0, b0 + 6)
        rank = (torch.rand(6, 5, generator=g) * 4).round() / 4           # ties on purpose
        preds.append([{'id': ids}, {'answer': tokens[b0:b0 + 6], 'rank': rank}])
    with tempfile.TemporaryDirectory() as d:
        TU.save_result(preds, ds, lambda data, idx: CU.to_sentence(idx, id2vocab), detok, d, 0, 3, 'cast_test')
        ans = open(os.path.join(d, 'result', 'cast_test_3.0.answer'), encoding='utf-8').read()
        run = open(os.path.join(d, 'result', 'cast_test_3.0.run'), encoding='utf-8').read()
    out = dict(tokens=rows, words=words_before, dedup=sents, answers=answers, extra=extra,
               save=dict(rank=[p[1]['rank'].tolist() for p in preds], answer_file=ans, run_file=run))
    with open(os.path.join(HERE, 'results_golden.json'), 'w') as f:
        json.dump(out, f)
    print('wrote results_golden.json:', len(rows), 'rows,', sum(len(a) != len(b) for a, b in zip(words_before, sents)), 'rows cut')


if __name__ == '__main__':
    main()
