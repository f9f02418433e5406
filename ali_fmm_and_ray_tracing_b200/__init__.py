"""ali_fmm_and_ray_tracing_b200 -- B200-native ALI-FMM travel-time fields and ray tracing.

The package holds only the accelerated hot path and its reference-facing boundary:

* ``csrc/``            hand-written sm_100a CUDA kernels + the C ABI (include/alifmm.h)
* ``_capi``            ctypes binding of libalifmm.so
* ``Anis_TTF_rays``    host-side mirror of the reference's ``ALI_FMM`` class

``from ali_fmm_and_ray_tracing_b200.Anis_TTF_rays import ALI_FMM`` (or the top-level
``Anis_TTF_rays`` shim of this repo) is a drop-in for the reference's module.
"""
from .Anis_TTF_rays import ALI_FMM, set_devices  # noqa: F401
from ._capi import AlifmmError, Context, device_count  # noqa: F401

__all__ = ["ALI_FMM", "set_devices", "AlifmmError", "Context", "device_count"]
