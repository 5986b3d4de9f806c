#!/usr/bin/env python
"""Golden vectors for the GLKS Mixturer (GLKS/Model.py:135-147) and Utils.copy_topk (common/Utils.py:170-178), produced by
the UNMODIFIED reference functions on seeded inputs (build container only; the .npz is committed).

    python tests/golden/make_glks_golden.py
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
from helpers import glks_inputs as inputs  # noqa: E402


def main():
    ns = refshim.load_reference()
    x = inputs()
    V = x['p_v'].size(1)
    mix = ns.glks.Mixturer(256)
    mix.load_state_dict({'linear1.weight': x['w'], 'linear1.bias': x['b']})
    with torch.no_grad():
        p = mix(x['state'], x['p_v'], x['p_k'], ns.utils.build_map(x['bmap'], max=V))
        tv, ti = ns.utils.topk(p.clone(), k=4)
        tv1, ti1 = ns.utils.topk(p.clone(), k=1)
        vm_onehot = ns.utils.build_map(x['vmap'], max=V)
        cv, ci = ns.utils.copy_topk(x['gen_ext'].clone(), vm_onehot, x['overlap'], k=5)
    out = dict(p=p, top4_v=tv, top4_i=ti, top1_v=tv1, top1_i=ti1, copy5_v=cv, copy5_i=ci)
    path = os.path.join(HERE, 'glks_mixturer.npz')
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
