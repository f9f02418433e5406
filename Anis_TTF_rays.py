"""Top-level drop-in for the reference's module name: ``from Anis_TTF_rays import ALI_FMM``
(as Weld_rays.py:2 and the example notebook do) resolves to the B200 implementation."""
from ali_fmm_and_ray_tracing_b200.Anis_TTF_rays import *  # noqa: F401,F403
from ali_fmm_and_ray_tracing_b200.Anis_TTF_rays import ALI_FMM, set_devices, tqdm_disable  # noqa: F401
