"""gstex_cuda_b200 -- B200 (sm_100a) native textured-2DGS rasteriser.

Drop-in for the ``gstex_cuda`` Python API of victor-rong/GStex_cuda: the sub-modules carry the reference's
names (``texture``, ``get_aabb_2d``, ``utils``, ``sh``, ``texture_sample``, ``timer``, ``cuda``,
``_torch_impl``) and the same public functions.  ``import gstex_cuda_b200 as gstex_cuda`` (or the alias
described in INTEGRATION.md) is all a caller changes.

The compute path is hand-written CUDA in ``libgstex_b200.so`` behind the C ABI of
``include/gstex_b200.h``; there is no CPU or library fallback.
"""
__version__ = "0.1.0"
