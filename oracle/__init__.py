"""CPU oracle for the CaSE_RG answer-decode hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package; the product (``case_rg_b200``) never does and has no CPU
fallback.

What it is: a torch-fp32 restatement of the reference's algorithm for the path (the reference is
pure Python/PyTorch, so torch on CPU *is* its arithmetic), written as plain
functions over a state_dict rather than nn.Modules:

* ``case_decoder.py``  - CaSETransformerSeqDecoder eval branch (CaSE/Model.py:38-48,50-63,91-125),
  both as the reference runs it (whole prefix recomputed every step, dense one-hot copy bmm) and
  as an incremental (KV-cached, index scatter) form used for large parity cases.
* ``generations.py``   - Generations.greedy / Generations.beam (common/Generations.py:66-220).
* ``gttp.py``          - GTTP BBCDecoder step + CopyGenerator (GTTP/Model.py:14-43,113-131,176-193).

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against outputs of the reference itself, run in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference through the shim of SURVEY.md §8c)
and committed under ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every fixture.
"""
