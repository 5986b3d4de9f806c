"""Output side: case_rg_b200.results against vectors produced by the unmodified reference
(tests/golden/make_results_golden.py), plus properties."""
import json
import os
import random

import torch

from case_rg_b200 import results as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'results_golden.json')


def _vocab(V=64):
    words = ['[PAD]', '[unused0]', '[unused1]'] + ['w%d' % i for i in range(3, V)]
    words[5], words[6], words[7] = '##ing', '[UNK]', '##s'
    return {i: w for i, w in enumerate(words)}


class _Dataset:
    def __init__(self, n, pool):
        self.samples = [dict(context_id=['c%d_%d' % (i, j) for j in range(i % 3)], query_id='q%d' % i,
                             passage_id=['p%d_%d' % (i, j) for j in range(1 + i % 2)],
                             passage_pool_id=['d%d_%d' % (i, j) for j in range(pool)]) for i in range(n)]

    def context_id(self, id): return self.samples[id]['context_id']
    def query_id(self, id): return self.samples[id]['query_id']
    def passage_id(self, id): return self.samples[id]['passage_id']
    def pool(self, id): return self.samples[id]['passage_pool_id']


def test_to_sentence_dedup_detokenise_match_reference_golden():
    z = json.load(open(GOLD))
    id2vocab = _vocab()
    tokens = torch.tensor(z['tokens'], dtype=torch.int64)
    sents = R.to_sentence(tokens, id2vocab)
    assert sents == z['words']
    assert R.to_sentence(z['tokens'], id2vocab) == z['words']          # nested lists work too
    R.remove_duplicate(sents)
    assert sents == z['dedup']
    detok = R.bert_detokenizer()
    assert [detok(s) for s in sents] == z['answers']
    assert R.answers_from_tokens(tokens, id2vocab) == z['answers']
    assert any(len(a) != len(b) for a, b in zip(z['words'], z['dedup']))


def test_remove_duplicate_other_n_match_reference_golden():
    z = json.load(open(GOLD))
    for case in z['extra']:
        t = [list(case['sent'])]
        R.remove_duplicate(t, case['n'])
        assert t[0] == case['out'], case


def test_remove_duplicate_is_a_fixed_point_and_quadratic_definition():
    """Against the literal definition (largest cut whose tail only has tokens seen before it), random cases."""
    rnd = random.Random(11)
    for _ in range(300):
        n = rnd.choice([1, 2, 3, 4])
        s = [rnd.randrange(0, rnd.choice([2, 3, 6])) for _ in range(rnd.randrange(0, 25))]
        want = list(s)
        while True:
            cut = -1
            if len(want) > n:
                for index in range(len(want) - n, 0, -1):
                    if all(e in want[:index] for e in want[index:]):
                        cut = index
                        break
            if cut < 0:
                break
            want = want[:cut]
        got = [list(s)]
        R.remove_duplicate(got, n)
        assert got[0] == want
        again = [list(got[0])]
        assert not R.remove_duplicate_once(again, n)


def test_save_result_files_match_reference_golden(tmp_path):
    z = json.load(open(GOLD))
    id2vocab = _vocab()
    tokens = torch.tensor(z['tokens'], dtype=torch.int64)
    ds = _Dataset(12, pool=5)
    preds = []
    for k, b0 in enumerate((0, 6)):
        preds.append([{'id': torch.arange(b0, b0 + 6)},
                      {'answer': tokens[b0:b0 + 6], 'rank': torch.tensor(z['save']['rank'][k], dtype=torch.float32)}])
    ap, rp = R.save_result(preds, ds, lambda data, idx: R.to_sentence(idx, id2vocab), R.bert_detokenizer(), str(tmp_path),
                           0, 3, 'cast_test')
    assert ap.endswith(os.path.join('result', 'cast_test_3.0.answer')) and rp.endswith('cast_test_3.0.run')
    assert open(ap, encoding='utf-8').read() == z['save']['answer_file']
    assert open(rp, encoding='utf-8').read() == z['save']['run_file']
    # answers only / ranks only
    ap2, rp2 = R.save_result([[p[0], {'answer': p[1]['answer']}] for p in preds], ds,
                             lambda data, idx: R.to_sentence(idx, id2vocab), R.bert_detokenizer(), str(tmp_path), 1, 0, 'x')
    assert ap2 is not None and rp2 is None


def test_empty_and_special_rows():
    id2vocab = _vocab()
    toks = torch.tensor([[0, 0, 0], [1, 2, 9], [2, 9, 9], [1, 0, 9]])
    assert R.to_sentence(toks, id2vocab) == [['[UNK]'], ['[UNK]'], ['[UNK]'], ['w9']]
    assert R.bert_detokenizer()(['play', '##ing', 'cat', '##s']) == 'playing cats'
