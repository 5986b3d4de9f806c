"""Reference arm: the UNMODIFIED PengjieRen/CaSE_RG sources, snapshotted into ``baseline/_ref`` (git-ignored, shipped to
the GPU box by gpurun) by ``baseline/make_ref.py`` and imported through ``baseline/refshim.py``.  Test / benchmark
infrastructure only: the product (``case_rg_b200``) never imports this package."""
