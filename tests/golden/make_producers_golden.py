#!/usr/bin/env python
"""Golden vectors for CaSE's pre-decode producers: the UNMODIFIED reference ``CaSE`` model runs
``forward(data, 'test')`` on seeded token ids and the tensors ``ResponseGeneration.action`` hands its decoder are captured
(CaSE/Model.py:247-251) together with the passage scores.  Build container only; the .npz is committed.

    python tests/golden/make_producers_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from baseline import refshim  # noqa: E402
from helpers import producers_case  # noqa: E402


def main():
    ns = refshim.load_reference()
    cfg, sd_prod, sd_dec, data, model = producers_case(ns)
    captured = {}
    dec = model.response_generation.decoder
    orig = dec.forward

    def spy(encode_memories, BOS, UNK, source_map, **kw):
        captured.update(mem_q=encode_memories[0], mem_p=encode_memories[1], answer_rep=kw['additional_decoder_feature'],
                        prior_q=kw['encode_weights'][0], prior_p=kw['encode_weights'][1])
        return orig(encode_memories, BOS, UNK, source_map, **kw)

    dec.forward = spy
    with torch.no_grad():
        res = model(dict(data), method='test')
        enc_p = model.query_encoder(data['passage'])[0][:, :, -1]
    out = dict(captured, rank=res['rank'], answer=res['answer'], enc_p=enc_p)
    path = os.path.join(HERE, 'case_producers.npz')
    np.savez_compressed(path, **{k: v.detach().numpy() for k, v in out.items()})
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
