"""Host-side mirror of the reference's sampler adapters (L4 of SURVEY.md):

    HMCModel + logdensity / logdensity_and_gradient    src/fitting/hmc_sample.jl:1-37
    MCMCModel callable                                  src/fitting/mcmc_sample.jl:1-24
    (new) MCMCModel.batch: W walkers per call through the batched-walker kernel (K6)

The adapters apply the O(T) variable transforms on the host exactly as the reference does and call the
device for everything that touches the template stack.
"""
from __future__ import annotations

import numpy as np

from .fitting import DeviceStack, device_stack


class HMCModel:
    """hmc_sample.jl:1-5.  ``composite`` is kept for signature parity (device scratch is used)."""

    def __init__(self, models, composite, data):
        self.models = device_stack(models, data)
        self.composite, self.data = composite, data

    def dimension(self):
        return self.models.shape[1]                                        # hmc_sample.jl:9

    def logdensity(self, theta):                                           # hmc_sample.jl:11-22
        theta = np.asarray(theta, dtype=np.float64)
        nl, _, _ = self.models.eval_fg(np.exp(theta), want_F=True, want_G=False)
        return -nl + theta.sum()

    __call__ = logdensity

    def logdensity_and_gradient(self, logx):                               # hmc_sample.jl:24-37
        logx = np.asarray(logx, dtype=np.float64)
        x = np.exp(logx)
        nl, G, _ = self.models.eval_fg(x, want_F=True, want_G=True)
        return -nl + logx.sum(), -G * x + 1

    def logdensity_and_gradient_batched(self, LOGX):
        """The same for C chains at once: LOGX is (T, C); one device pass (sfh_eval_fg_batched) serves all chains."""
        LOGX = np.asarray(LOGX, dtype=np.float64)
        Xn = np.exp(LOGX)
        nl, G = self.models.eval_fg_batched(Xn)
        return -nl + LOGX.sum(axis=0), -G * Xn + 1


class MCMCModel:
    """mcmc_sample.jl:1-24: log-likelihood only; negative coefficients -> -Inf."""

    def __init__(self, models, data):
        self.models = device_stack(models, data)
        self.data = data

    def dimension(self):
        return self.models.shape[1]                                        # mcmc_sample.jl:9

    def __call__(self, theta):                                             # mcmc_sample.jl:12-23
        return float(self.models.eval_logl_batched(np.asarray(theta, dtype=np.float64)[:, None])[0])

    logdensity = __call__

    def batch(self, X):
        """X: (T, W) -- one column per walker.  Returns W log-likelihoods from ONE device pass."""
        return self.models.eval_logl_batched(X)
