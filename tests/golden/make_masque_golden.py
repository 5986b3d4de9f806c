#!/usr/bin/env python
"""Golden fixture for the Masque decoder face: runs the UNMODIFIED reference
``MasqueTransformerSeqDecoder`` (Masque/Model.py:13-119) in eval mode on seeded weights / inputs.

    python tests/golden/make_masque_golden.py      (build container only: imports /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import install_reference, syn, build_map   # noqa: E402  (installs the import shim)

import Masque.Model as ref_masque                             # noqa: E402

V, H = 1000, 256


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    cfg = dict(wseed=31, iseed=41, B=4, Lq=10, NP=4, Lp=12, T=10, peaked=0.35, pad_boost=14.0, gate=2.0)
    sd = syn.make_masque_decoder_state(cfg['wseed'], V, H, peaked=cfg['peaked'], boost={0: cfg['pad_boost']},
                                       gen_gate_bias=cfg['gate'])
    inp = syn.make_case_inputs(cfg['iseed'], cfg['B'], cfg['Lq'], cfg['NP'], cfg['Lp'], V, H)
    dec = ref_masque.MasqueTransformerSeqDecoder(2, 4, 8, V, H)
    dec.load_state_dict(sd, strict=True)
    dec.eval()
    out = {}
    with torch.no_grad():
        for tag, weights in (('w', inp.encode_weights), ('now', None)):
            dec_out, gen, dist, toks = dec(inp.encode_memories, syn.BOS, syn.UNK, build_map(inp.source_map, max=V),
                                           encode_masks=inp.encode_masks, encode_weights=weights,
                                           max_target_length=cfg['T'])
            print('masque greedy', tag, toks.tolist())
            out.update({f'tokens_{tag}': toks.numpy(), f'dist_{tag}': dist.numpy(), f'dec_out_{tag}': dec_out.numpy()})
    np.savez_compressed(os.path.join(HERE, 'masque_module_greedy.npz'), wsum=syn.state_checksum(sd),
                        cfg=np.array(list(cfg.values()), dtype=np.float64), cfg_keys=np.array(list(cfg.keys())), **out)


if __name__ == '__main__':
    main()
