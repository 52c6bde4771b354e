"""Inference half of the reference's `diff_params.edm.EDM` (edm.py:7-148): schedule, stochasticity, preconditioning.

Select with `diff_params.callable: "audio-inpainting-diffusion_b200.edm.EDM"`; the reference's own EDM object works
equally well with this package's Sampler (only the attributes/methods below are used).  Training-side
methods (loss_fn, sample_ptrain*, edm.py:67-85,150-193) are out of scope for this path.
"""
import torch

from .config import cfg_get


class EDM:
    def __init__(self, args):
        self.args = args
        g = lambda k: cfg_get(args, "diff_params." + k)
        self.sigma_min, self.sigma_max = g("sigma_min"), g("sigma_max")
        self.P_mean, self.P_std = cfg_get(args, "diff_params.P_mean", -1.2), cfg_get(args, "diff_params.P_std", 1.2)
        self.ro, self.ro_train = g("ro"), cfg_get(args, "diff_params.ro_train", g("ro"))
        self.sigma_data = g("sigma_data")
        self.Schurn, self.Stmin, self.Stmax, self.Snoise = g("Schurn"), g("Stmin"), g("Stmax"), g("Snoise")
        if cfg_get(args, "diff_params.aweighting.use_aweighting", False):
            raise NotImplementedError("A-weighting is a training-loss option (edm.py:33-34), not on the sampling path")

    def get_gamma(self, t):
        """edm.py:38-53 (N = len(t))."""
        N = t.shape[0]
        gamma = torch.zeros(t.shape).to(t.device)
        indexes = torch.logical_and(t > self.Stmin, t < self.Stmax)
        gamma[indexes] = gamma[indexes] + torch.min(torch.Tensor([self.Schurn / N, 2 ** (1 / 2) - 1]))
        return gamma

    def create_schedule(self, nb_steps):
        """edm.py:55-64"""
        i = torch.arange(0, nb_steps + 1)
        t = (self.sigma_max ** (1 / self.ro) + i / (nb_steps - 1) * (self.sigma_min ** (1 / self.ro) - self.sigma_max ** (1 / self.ro))) ** self.ro
        t[-1] = 0
        return t

    def sample_prior(self, shape, sigma):
        """edm.py:87-95: noise is drawn on the CPU generator, then moved."""
        return torch.randn(shape).to(sigma.device) * sigma

    def cskip(self, sigma):
        return self.sigma_data ** 2 * (sigma ** 2 + self.sigma_data ** 2) ** -1

    def cout(self, sigma):
        return sigma * self.sigma_data * (self.sigma_data ** 2 + sigma ** 2) ** (-0.5)

    def cin(self, sigma):
        return (self.sigma_data ** 2 + sigma ** 2) ** (-0.5)

    def cnoise(self, sigma):
        return (1 / 4) * torch.log(sigma)

    def denoiser(self, xn, net, sigma):
        """edm.py:133-148.  Uses the fused device entry point when `net` provides one."""
        if len(sigma.shape) == 1:
            sigma = sigma.unsqueeze(-1)
        cskip, cout, cin, cnoise = self.cskip(sigma), self.cout(sigma), self.cin(sigma), self.cnoise(sigma)
        if hasattr(net, "denoise_fused") and sigma.numel() == 1:
            # the preconditioning is fused into the CUDA call, on the differentiable path too (the VJP carries the same scales)
            return net.denoise_fused(xn, cnoise, float(cin), float(cout), float(cskip))
        return cskip * xn + cout * net(cin * xn, cnoise)
