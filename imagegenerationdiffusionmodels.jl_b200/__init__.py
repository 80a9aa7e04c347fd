"""B200-native DDPM hot path behind ImageGenerationDiffusionModels.jl's API.

Sub-modules: ``capi`` (ctypes binding of libddpm.so), ``api`` (host mirror of the reference's
public functions), ``tables`` (schedule/embedding constants), ``bson_io`` (BSON.jl checkpoints),
``dist`` (one-process-per-GPU sharding and communicator bootstrap), ``build`` (nvcc recipe).

The directory name contains a dot, so import it through the repo-root shim::

    import igdm_b200                     # == this package
    from igdm_b200 import api, capi
"""
__all__ = ["api", "bson_io", "build", "capi", "dist", "tables"]
