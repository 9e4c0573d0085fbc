"""ctxnerf -- B200-native (sm_100a) NeRF ray-march path behind the
``src/run_nerf_helpers.py`` API of ConTEXTure-NeRF.

    from ctxnerf import run_nerf_helpers            # the drop-in module
    from ctxnerf.run_nerf_helpers import *          # what trainer.py:33 does

Everything numeric runs in ``libctxnerf.so`` (C-ABI in include/ctxnerf.h);
importing this package does not load the library, the first op does, and it
raises if the library has not been built.
"""
__version__ = "0.1.0"
