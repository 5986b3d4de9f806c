"""case_rg_b200 - B200-native answer-decode path for PengjieRen/CaSE_RG.

Host-side mirror of the reference interface for this path:
  FastCaSEDecoder / install_fast_decoder ...... CaSETransformerSeqDecoder face (CaSE/Model.py:13-125)
  generations.greedy / beam, FastCaSE, FastGTTP  EncDecModel + Generations face (common/Generations.py)
  engine.* ..................................... device buffers + the C-ABI step calls
  distributed.* ................................ rank sharding + the final gather
  results.* .................................... to_sentence / remove_duplicate / save_result (.answer, .run files)
  synthetic.* .................................. seeded CAsT-shaped inputs / random-init checkpoints
The kernels live in csrc/ and are reached only through libcase_b200.so (include/case_b200.h).
"""
__all__ = ['synthetic']
