#!/usr/bin/env python
"""`CaSE/Run.py --mode test` with the B200 path installed (CaSE/Run.py:54-62, common/CumulativeTrainer.py:134-156).

Builds the UNMODIFIED reference `CaSE` model from the snapshot in baseline/_ref (random-init weights: there are no
checkpoints offline), installs the device path into it - `--install decoder`: only the answer decoder,
`--install model`: the pre-decode producers too (8f N1) - and runs the predict loop over synthetic CAsT-shaped batches:
`model(data, method='test')` per batch, answers turned into strings with the reference's own post-processing.

    python examples/case_test_mode.py [--install decoder|model] [--batches 4] [--batch-size 64] [--beam 4]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from baseline import refshim                                    # noqa: E402
from case_rg_b200 import synthetic as syn                       # noqa: E402
from case_rg_b200.decoder import install_fast_decoder, install_fast_model    # noqa: E402
from case_rg_b200.results import answers_from_tokens            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--install', default='model', choices=['decoder', 'model'])
    ap.add_argument('--batches', type=int, default=4)
    ap.add_argument('--batch-size', type=int, default=64)
    ap.add_argument('--beam', type=int, default=4)
    ap.add_argument('--passages', type=int, default=10)
    ap.add_argument('--passage-len', type=int, default=256)
    ap.add_argument('--max-target-length', type=int, default=40)
    args = ap.parse_args()
    if refshim.reference_root() is None:
        sys.exit('reference snapshot missing: run `python baseline/make_ref.py` where /root/reference exists')
    ns = refshim.load_reference()
    V, H, dev = syn.BERT_VOCAB, 256, 'cuda'
    T = args.max_target_length
    model = refshim.reference_case_model(ns, V, T, decoder_sd=syn.make_case_decoder_state(1, V, H, peaked=0.3)).to(dev).eval()
    (install_fast_model if args.install == 'model' else install_fast_decoder)(model, beam_width=args.beam, dtype='bf16')
    _, id2vocab = syn.make_vocab(V)
    batches = []
    for b in range(args.batches):
        inp = syn.make_case_inputs(100 + b, args.batch_size, 60, args.passages, args.passage_len, V, H)
        batches.append({'id': inp.ids, 'query': inp.query.pin_memory(), 'passage': inp.passage.pin_memory(),
                        'source_map': inp.source_map.pin_memory()})
    outs, toks = [], 0
    with torch.no_grad():
        for rep in range(2):                                     # first pass: builds engines and captures the decode graph
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            outs = []
            for data in batches:                                 # CumulativeTrainer.predict: one forward per batch
                d = {k: v.to(dev, non_blocking=True) for k, v in data.items()}
                outs.append(model(d, method='test'))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
    for o in outs:
        toks += int((o['answer'] != 0).sum())
    print(f'{args.install}: {len(batches)} batches of {args.batch_size} queries in {dt * 1e3:.1f} ms '
          f'({dt * 1e3 / len(batches):.2f} ms per batch, {toks / dt:.0f} answer tokens/s)')
    first = answers_from_tokens(outs[0]['answer'][:2].cpu(), id2vocab)
    for i, s in enumerate(first):
        print(f'  query {i}: rank of passages {outs[0]["rank"][i].argsort(descending=True)[:3].tolist()}  answer: {s[:100]}')


if __name__ == '__main__':
    main()
