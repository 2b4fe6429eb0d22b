"""gstools_b200 -- B200-native (sm_100a CUDA) backend for GSTools' randomisation-method summation.

One hot path and nothing else: the native ``summate`` / ``summate_incompr`` functions that
``gstools.field.generator.RandMeth`` / ``IncomprRandMeth`` call through the backend switch
(reference: src/gstools/field/generator.py:42-64).  Use it

* directly -- ``gstools_b200.summate(cov_samples, z_1, z_2, pos, num_threads=None)`` with the
  reference's signature, numpy in / numpy out or CUDA torch tensors in / out;
* as a drop-in third backend of an unmodified gstools -- ``gstools_b200.enable()``; then
  ``gs.SRF(model)(pos)`` and ``gs.CondSRF(krige)(...)`` run on the GPU unchanged.

The compute lives in ``libgsb200.so`` (C ABI in ``include/gsb200.h``); there is no CPU fallback.
"""

from . import _build, _lib
from ._lib import (GSB200Error, device_count, get_counter, kernel_times, measure_fp64_peak,
                   release_memory, set_option)
from .backend import (
    calc_field_krige,
    calc_field_krige_and_variance,
    cov_model_spec,
    krige_evaluate,
    sample_radii_mcmc,
    sample_modes_batch,
    get_device,
    cond_scaling,
    make_epilogue,
    make_point_epilogue,
    scale_shift_,
    set_device,
    summate,
    summate_fourier,
    summate_fourier_structured,
    summate_incompr,
    summate_incompr_structured,
    summate_structured,
    Plan,
    use_devices,
    current_plan,
)
from .plugin import disable, enable, ensemble, is_enabled, prewarm, unfused_methods

__version__ = "0.1.0"

__all__ = [
    "summate",
    "summate_incompr",
    "summate_structured",
    "summate_incompr_structured",
    "summate_fourier",
    "summate_fourier_structured",
    "calc_field_krige_and_variance",
    "calc_field_krige",
    "krige_evaluate",
    "cov_model_spec",
    "sample_radii_mcmc",
    "sample_modes_batch",
    "scale_shift_",
    "make_epilogue",
    "make_point_epilogue",
    "cond_scaling",
    "enable",
    "disable",
    "is_enabled",
    "prewarm",
    "ensemble",
    "unfused_methods",
    "set_device",
    "get_device",
    "Plan",
    "use_devices",
    "current_plan",
    "device_count",
    "get_counter",
    "set_option",
    "release_memory",
    "measure_fp64_peak",
    "kernel_times",
    "GSB200Error",
    "build",
]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile ``libgsb200.so`` for sm_100a (nvcc cross-compiles without a GPU)."""
    return _build.build(force=force, verbose=verbose)
