#!/usr/bin/env python
"""Turns the per-case JSON files the search-parity tests write (tests/test_gpu_search_parity.py ->
gpurun_out/parity_<case>.json) into the table committed as profiles/r2_parity.md.
usage: python profiles/micro/parity_summary.py [gpurun_out] > profiles/r2_parity.md"""
import glob
import json
import os
import sys


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out'
    rows = []
    for f in sorted(glob.glob(os.path.join(src, 'parity_*.json'))):
        d = json.load(open(f))
        name = os.path.basename(f)[len('parity_'):-len('.json')]
        if 'rel_err' in d:
            if isinstance(d['rel_err'], dict):         # producers at the full C2 size: max relative error per output
                txt = ', '.join(f'{k} {v:.1e}' for k, v in d['rel_err'].items())
                rows.append((name, 'pre-decode producers, rel. error vs the fp32 oracle', txt, '', '', '', ''))
            else:
                rows.append((name, 'logits of step 0', f"rel. error {d['rel_err']:.2e}", '', '', '', ''))
            continue
        n = d.get('rows', d.get('queries'))
        kind = 'greedy rows' if 'rows' in d else 'beam queries'
        agree = d.get('decision_agreement')
        if agree is None and d.get('decisions'):
            agree = 1.0 - d['flipped'] / d['decisions']
        rows.append((name, f'{n} {kind}', str(d['identical']), str(d['near_tie']), str(d['miss']),
                     f"{d['max_gap_nats']:.4f} / {d['tol_nats']:.4f}", '' if agree is None else f'{agree:.4f}'))
    print('# Search-path parity at the BASELINE shapes (round 2)\n')
    print('CUDA path against the oracle evaluated on the same GPU in strict fp32 (TF32 off), same seeded peaked weights and')
    print('inputs, through the C ABI (`tests/test_gpu_search_parity.py`).  A differing answer is scored with the oracle: greedy -')
    print('the oracle log-probability gap between its own choice and the CUDA token at the first differing step; beam - the')
    print('difference of the two answers\' keys (cumulative cost / length) under the oracle model.  `near tie` = gap within the')
    print('tolerance `rel x max|logit|` nats (rel = 1e-4 fp32, 2e-2 bf16 storage, BASELINE.json north_star), `miss` = beyond it.')
    print('fp32 storage must be identical.\n')
    print('| case | size | identical | near tie | miss | max gap / tolerance (nats) | per-decision agreement |')
    print('|---|---|---|---|---|---|---|')
    for r in rows:
        print('| ' + ' | '.join(r) + ' |')


if __name__ == '__main__':
    main()
