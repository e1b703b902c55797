"""genmmrec_b200 -- B200-native hot path of GenMMRec (sparse graph propagation + full-sort
evaluation) behind the reference's GeneralRecommender / Trainer.evaluate API.

Layout: ``csrc/`` hand-written sm_100a CUDA behind the C ABI of ``include/gmr.h`` (built to
``libgmr.so`` by ``build.py``), ``_lib.py`` the ctypes binding, ``ops.py`` the tensor-level
operators, ``graph.py`` the CSR containers and vectorised graph builders, and
``common/ models/ utils/`` the host-side mirror of the reference's plugin interface.
There is no CPU fallback: every operator raises if ``libgmr.so`` cannot be loaded.
"""
__version__ = "0.1.0"
