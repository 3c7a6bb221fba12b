"""sfh_b200 -- B200-native fitting hot path of StarFormationHistories.jl.

The directory is named ``starformationhistories.jl_b200`` (not importable by name because of the dot);
import it through the ``sfh_b200`` shim at the repository root::

    import sfh_b200 as sfh
    stack = sfh.DeviceStack(models, data)               # upload once per fit
    nlogL = sfh.fg_(True, G, coeffs, stack, data)       # fg!  (src/fitting/solvers.jl:20-38)

Importing this package loads ``libsfhcuda.so`` and fails loudly if it is missing.
"""
from . import _lib
from ._lib import SFHError, device_count
from .fitting import (DeviceStack, DeviceStackGroup, clear_cache, composite_, device_stack, fg_ as fg_flat_, grad_loglikelihood,
                      grad_loglikelihood_, loglikelihood, stack_models)
from .hierarchical import (GaussianDispersion, HierarchicalOptimizer, LinearAMR, LogarithmicAMR, PowerLawMZR,
                           calculate_coeffs, exptransform, fg_ as fg_hier_, logtransform, nparams, MH_from_Z, dMH_dZ, Z_from_MH, dZ_dMH,
                           X_from_Z, Y_from_Z)
from .sampling import HMCModel, MCMCModel
from . import io, sharding, solvers
from .io import SFHFile, load_result, read_arrays, save_result, write_arrays
from .solvers import (calculate_cum_sfr, construct_x0, construct_x0_mdf, cum_sfr_quantiles, fixed_amr, truncate_relweights, fit_sfh, fit_templates, fit_templates_fast, fit_templates_lbfgsb,
                      hmc_sample, mcmc_sample, mdf_amr, rand_result, renormalize_x0, result_median, result_mode, result_std, sample_sfh, tau, tau_interp, tsample_sfh)
from . import templates
from .templates import bin_cmd, bin_cmd_smooth, build_template_stack, partial_cmd, partial_cmd_smooth, template_points
from .sharding import allreduce_fg, guard_neg_logl, init_library_comm, shard_rows


def fg_(F, G, *args):
    """``fg!`` with the reference's two method families (multiple dispatch on the third argument):

    ``fg_(F, G, coeffs, models, data[, composite])``                                    solvers.jl:20-38
    ``fg_(F, G, MHmodel0, dispmodel0, variables, models, data, composite, logAge, MH)``  mzr.jl:84 / amr.jl:78
    """
    if args and hasattr(args[0], "kind") and hasattr(args[0], "free_params"):
        return fg_hier_(F, G, *args)
    return fg_flat_(F, G, *args)


__all__ = ["DeviceStack", "DeviceStackGroup", "SFHError", "device_count", "stack_models", "composite_", "loglikelihood",
           "grad_loglikelihood", "grad_loglikelihood_", "fg_", "calculate_coeffs", "PowerLawMZR", "LinearAMR",
           "LogarithmicAMR", "GaussianDispersion", "HierarchicalOptimizer", "HMCModel", "MCMCModel", "nparams",
           "exptransform", "logtransform", "clear_cache", "device_stack", "shard_rows", "allreduce_fg", "guard_neg_logl",
           "init_library_comm", "fit_templates_lbfgsb", "fit_templates", "fit_templates_fast", "fit_sfh", "mcmc_sample",
           "hmc_sample", "renormalize_x0", "mdf_amr", "calculate_cum_sfr", "cum_sfr_quantiles", "tau", "tau_interp", "rand_result", "result_mode", "result_median", "result_std", "construct_x0", "bin_cmd_smooth", "partial_cmd_smooth", "build_template_stack",
           "template_points", "templates", "sample_sfh", "tsample_sfh", "fixed_amr", "truncate_relweights", "construct_x0_mdf", "bin_cmd", "partial_cmd",
           "MH_from_Z", "dMH_dZ", "Z_from_MH", "dZ_dMH", "X_from_Z", "Y_from_Z", "io", "SFHFile", "write_arrays", "read_arrays", "save_result", "load_result"]
