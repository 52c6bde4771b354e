"""ORACLE (test infrastructure, not product): plain-PyTorch fp32 CPU restatement of the denoiser path.

Follows, function by function, the reference files (read-only at /root/reference, never copied):
  networks/unet_cqt_oct_with_projattention_adaLN_2.py   (abbreviated unet.py below)
  diff_params/edm.py                                    (edm.py)
  testing/edm_sampler_inpainting.py                     (sampler.py)

It is a *functional* restatement driven by a reference-schema state dict, so that it runs on the GPU
box where /root/reference does not exist.  It is pinned by tests/test_oracle_vs_reference.py (run in the
build container, where the real reference modules are importable) and by the golden vectors under
tests/golden/ that tests/golden/make_golden.py produced from the reference's own code.  The CQT it
calls is oracle/cqt_oracle.py (parity unpinned there -- see that file's header).

Only tests/, bench.py's cpu_baseline / reference arm and __graft_entry__.smoke() may import this.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from cqt_oracle import CQT_nsgt

SQRT2 = 2 ** 0.5

# unet.py:513-515  ('cubic' resampling taps)
CUBIC = [-0.01171875, -0.03515625, 0.11328125, 0.43359375, 0.43359375, 0.11328125, -0.03515625, -0.01171875]


def group_norm(x, gamma, groups=8, eps=1e-7):
    """unet.py:147-163  x / (unbiased std over (C/g, F, T) + eps) * gamma; no centring."""
    B, C, Fd, T = x.shape
    std = x.reshape(B, groups, -1).std(-1, keepdim=True)
    y = (x.reshape(B, groups, -1) / (std + eps)).reshape(B, C, Fd, T)
    return y * gamma


def linear(sd, name, e):
    """unet.py:36-40"""
    return e @ sd[name + ".weight"].t() + sd[name + ".bias"]


def conv(sd, name, x, dilation=1):
    """unet.py:79-88  bias-free conv, zero 'same' padding, dilation along F."""
    return F.conv2d(x, sd[name + ".weight"], padding="same", dilation=dilation)


def embedding(sd, c_noise):
    """unet.py:184-211  RFF (sin, cos of 2*pi*s*f) then 3 x (Linear, ReLU)."""
    table = 2 * np.pi * c_noise * sd["embedding.RFF_freq"]
    e = torch.cat([torch.sin(table), torch.cos(table)], dim=1)
    for i in range(3):
        e = F.relu(linear(sd, f"embedding.MLP.{i}", e))
    return e


def down_t(x):
    """unet.py:549-580 (down): reflect pad 3|3, 8-tap FIR, stride 2 along T."""
    B, C, Fd, T = x.shape
    k = torch.tensor(CUBIC, dtype=x.dtype).view(1, 1, 8)
    y = F.conv1d(F.pad(x.reshape(-1, 1, T), (3, 3), mode="reflect"), k, stride=2)
    return y.reshape(B, C, Fd, -1)


def up_t(x):
    """unet.py:549-580 (up): reflect pad 2|2, transposed 8-tap FIR, stride 2, padding 7 -> 2T."""
    B, C, Fd, T = x.shape
    k = torch.tensor(CUBIC, dtype=x.dtype).view(1, 1, 8)
    y = F.conv_transpose1d(F.pad(x.reshape(-1, 1, T), (2, 2), mode="reflect"), k, stride=2, padding=7)
    return y.reshape(B, C, Fd, -1)


def time_attention(sd, p, x, heads):
    """unet.py:338-380"""
    B, _, Fd, T = x.shape
    h = conv(sd, p + ".proj_in", x)  # [B,heads,F,T]
    hf = h.reshape(B, heads * Fd, T)
    v = h.permute(0, 1, 3, 2)  # [B,heads,T,F] identity values
    qk = F.conv1d(hf, sd[p + ".qk.weight"])  # [B,2*heads*F,T]
    qk = qk.reshape(B, heads, 2 * Fd, T).permute(0, 1, 3, 2)
    q, k = qk[..., :Fd], qk[..., Fd:]
    sim = torch.matmul(q, k.transpose(-1, -2)) * (float(Fd) ** -0.5)
    out = torch.matmul(sim.softmax(dim=-1), v).permute(0, 1, 3, 2)  # [B,heads,F,T]
    return conv(sd, p + ".proj_out", out)


def resnet_block(sd, p, x_in, emb, *, dim, dim_out, num_dils, k1x1=False, after=False, attention=False, heads=8):
    """unet.py:452-493"""
    N = dim if after else dim_out
    x = conv(sd, p + ".proj_in", x_in) if dim != N else x_in
    if attention:
        i_x = x
        g = linear(sd, p + ".affine2", emb)[:, :, None, None]
        s = linear(sd, p + ".gate2", emb)[:, :, None, None]
        x = group_norm(x, sd[p + ".norm2.gamma"]) * (g + 1)
        x = time_attention(sd, p + ".attn_block", x, heads) * s
        x = (x + i_x) / SQRT2
    for i in range(num_dils):
        x0 = x
        g = linear(sd, f"{p}.affine.{i}", emb)[:, :, None, None]
        s = linear(sd, f"{p}.gate.{i}", emb)[:, :, None, None]
        x = group_norm(x, sd[f"{p}.norm.{i}.gamma"]) * (g + 1)
        x = (x0 + conv(sd, f"{p}.H.{i}", F.gelu(x), dilation=1 if k1x1 else (2 ** i, 1)) * s) / SQRT2
    if after and N != dim_out:
        x = conv(sd, p + ".proj_out", x)
    res = conv(sd, p + ".res_conv", x_in) if dim != dim_out else x_in
    return (x + res) / SQRT2


class UnetOracle:
    """unet.py:583-845 as a function of (state dict, config)."""

    def __init__(self, cfg, sd):
        self.cfg, self.sd = cfg, sd
        win = ("kaiser", cfg["beta"]) if cfg["window"] == "kaiser" else cfg["window"]
        self.CQTransform = CQT_nsgt(cfg["num_octs"], cfg["bins_per_oct"], mode="oct", window=win,
                                    fs=cfg["sample_rate"], audio_len=cfg["audio_len"], dtype=torch.float32)

    def taps(self):
        return None

    @torch.no_grad()
    def __call__(self, inputs, sigma, probe=None):
        return self._forward(inputs, sigma, probe)

    def differentiable(self, inputs, sigma):
        """The same forward with autograd recording (reference for the input gradient, sampler.py:59-78)."""
        with torch.enable_grad():
            return self._forward(inputs, sigma, None)

    def _forward(self, inputs, sigma, probe=None):
        cfg, sd = self.cfg, self.sd
        Ns, nd, att, bins, no = cfg["Ns"], cfg["num_dils"], cfg["attention_layers"], cfg["bins_per_oct"], cfg["num_octs"]
        emb = embedding(sd, sigma)
        X_list = self.CQTransform.fwd(inputs.unsqueeze(1))
        outs = [None] * no
        hs = []
        X = pyr = None
        for i in range(no):
            C = torch.view_as_real(X_list[-1 - i].squeeze(1)).permute(0, 3, 1, 2).contiguous()  # [B,2,bins,T_i]
            din = Ns[i] if i == 0 else Ns[i - 1]
            C2 = resnet_block(sd, f"downs.{i}.0", C, emb, dim=2, dim_out=din, num_dils=1, k1x1=True)
            if i == 0:
                X, pyr = C2, down_t(C)
            elif i < no - 1:
                pyr = torch.cat((down_t(C), down_t(pyr)), dim=2)
                X = torch.cat((C2, X), dim=2)
            else:
                pyr = torch.cat((C, pyr), dim=2)
                X = torch.cat((C2, X), dim=2)
            X = resnet_block(sd, f"downs.{i}.2", X, emb, dim=din, dim_out=Ns[i], num_dils=nd[i], attention=bool(att[i]))
            if probe is not None:
                probe[f"enc{i}"] = X
            hs.append(X)
            if i < no - 1:
                X = down_t(X)
            X = (X + conv(sd, f"downs.{i}.1", pyr)) / SQRT2
        X = resnet_block(sd, "middle.0.1", X, emb, dim=Ns[-1], dim_out=Ns[-1], num_dils=nd[-1], attention=bool(att[-1]))
        Xout = resnet_block(sd, "middle.0.0", X, emb, dim=Ns[-1], dim_out=2, num_dils=1, k1x1=True, after=True)
        if probe is not None:
            probe["mid"] = X
        for i in range(no):
            j = no - 1 - i
            dout = Ns[j] if j == 0 else Ns[j - 1]
            X = torch.cat((X, hs.pop()), dim=1)
            X = resnet_block(sd, f"ups.{i}.1", X, emb, dim=2 * Ns[j], dim_out=dout, num_dils=nd[j], attention=bool(att[j]))
            Xout = (Xout + resnet_block(sd, f"ups.{i}.0", X, emb, dim=dout, dim_out=2, num_dils=1, k1x1=True, after=True)) / SQRT2
            if probe is not None:
                probe[f"dec{i}"] = X
            X = X[:, :, bins:, :]
            Out, Xout = Xout[:, :, :bins, :], Xout[:, :, bins:, :]
            outs[i] = torch.view_as_complex(Out.permute(0, 2, 3, 1).contiguous()).unsqueeze(1)
            if j > 0:
                X, Xout = up_t(X), up_t(Xout)
        pred = self.CQTransform.bwd(outs).squeeze(1)[:, : inputs.shape[-1]]
        assert pred.shape == inputs.shape
        return pred


# ---- edm.py ---------------------------------------------------------------------------------------
class EDMOracle:
    """edm.py:55-64, 38-53, 97-148 restated on plain floats/tensors."""

    def __init__(self, sigma_data=0.063, sigma_min=1e-4, sigma_max=1.0, ro=13, Schurn=10, Snoise=1.0, Stmin=0, Stmax=50):
        self.sigma_data, self.sigma_min, self.sigma_max, self.ro = sigma_data, sigma_min, sigma_max, ro
        self.Schurn, self.Snoise, self.Stmin, self.Stmax = Schurn, Snoise, Stmin, Stmax

    def create_schedule(self, nb_steps):  # edm.py:55-64
        i = torch.arange(0, nb_steps + 1)
        t = (self.sigma_max ** (1 / self.ro) + i / (nb_steps - 1) * (self.sigma_min ** (1 / self.ro) - self.sigma_max ** (1 / self.ro))) ** self.ro
        t[-1] = 0
        return t

    def get_gamma(self, t):  # edm.py:38-53 (N = len(t), not nb_steps)
        N = t.shape[0]
        gamma = torch.zeros(t.shape)
        idx = torch.logical_and(t > self.Stmin, t < self.Stmax)
        gamma[idx] = gamma[idx] + min(self.Schurn / N, 2 ** 0.5 - 1)
        return gamma

    def denoiser(self, xn, net, sigma):  # edm.py:133-148
        if sigma.dim() == 1:
            sigma = sigma.unsqueeze(-1)
        sd = self.sigma_data
        cskip = sd ** 2 * (sigma ** 2 + sd ** 2) ** -1
        cout = sigma * sd * (sd ** 2 + sigma ** 2) ** (-0.5)
        cin = (sd ** 2 + sigma ** 2) ** (-0.5)
        cnoise = (1 / 4) * torch.log(sigma)
        return cskip * xn + cout * net(cin * xn, cnoise)


def smooth_mask(mask, size):
    """sampler.py:302-325 vectorised: half-Hann ramps of `size` samples on both sides of every gap of row 0."""
    hann = torch.hann_window(size * 2)
    left, right = hann[:size], hann[size:]
    m = mask[0]
    new = m.clone()
    prev = torch.cat((torch.ones(1, dtype=m.dtype), m[:-1]))
    for i in torch.nonzero(m != prev)[:, 0].tolist():
        if m[i] == 0:
            new[i - size:i] = right
        if m[i] == 1:
            new[i:i + size] = left
    return new.unsqueeze(0).expand(mask.shape[0], -1)


def spectral_mask(x, mask, n_fft=1024, hop=256, win_length=1024):
    """sampler.py:271-290 (= tester_inpainting.py:298-322): zero-pad to a multiple of n_fft (a whole n_fft when it already is
    one), torch.stft with a periodic Hann window (centre = True, reflect), multiply by the real [n_fft/2+1, frames] mask,
    torch.istft, crop to the input length."""
    window = torch.hann_window(win_length)
    L = x.shape[-1]
    xp = F.pad(x, (0, n_fft - L % n_fft), mode="constant", value=0)
    X = torch.stft(xp, n_fft, hop, win_length, window, return_complex=True)
    X = X * mask.unsqueeze(0)
    return torch.istft(X, n_fft, hop, win_length, window, return_complex=False)[..., 0:L]


def spectral_projection(y, mask, **stft):
    """sampler.py:361: proj_convex_set of the spectrogram-inpainting mode, x -> y + x - S(x)."""
    return lambda x: y + x - spectral_mask(x, mask, **stft)


def guided_estimate(net, edm, x, ti, y, degradation, xi, audio_len, norm_ord=2, hpf=None):
    """sampler.py:55-95 (reconstruction guidance, batch 1 as in the reference -- its autograd.grad call needs a scalar):
    x_hat - s * d||y - degradation(x_hat)|| / dx with s = t_i * xi / (||grad|| / sqrt(audio_len) + 1e-6)."""
    x = x.detach().requires_grad_()
    with torch.enable_grad():
        xh = edm.denoiser(x, net, ti.reshape(1, 1))
        if hpf is not None:
            xh = hpf(xh)
        norm = torch.linalg.norm(y - degradation(xh), dim=1, ord=norm_ord)
        g = torch.autograd.grad(outputs=norm, inputs=x)[0]
    s = ti * xi / (torch.linalg.norm(g) / audio_len ** 0.5 + 1e-6)
    return xh.detach() - s * g


def sample_oracle(net, edm, shape, noises, nb_steps=35, order=2, y=None, mask_s=None, hpf=None, project=None, guidance=None):
    """sampler.py:178-262 with the xi=0 replacement branch (141-147) or the unconditional branch (116-125).
    `project` (a callable) replaces the time-domain projection `mask_s*y + (1-mask_s)*x` by another proj_convex_set.
    `guidance` = dict(xi, degradation, audio_len[, norm_ord]) selects the xi > 0 branch (sampler.py:133-135, 55-113): the
    guided estimate first, then the projection if one is given (data_consistency).

    `noises` is an iterator of pre-drawn standard normal tensors consumed in the reference's order:
    first the prior (edm.py:94), then one per stochastic step (sampler.py:212).
    """
    t = edm.create_schedule(nb_steps)
    gamma = edm.get_gamma(t)
    noises = iter(noises)
    x = next(noises) * t[0]

    def score(x, ti):
        if guidance is not None:
            xh = guided_estimate(net, edm, x, ti, y, guidance["degradation"], guidance["xi"], guidance["audio_len"],
                                 guidance.get("norm_ord", 2), hpf)
            if project is not None:
                xh = project(xh)
            elif mask_s is not None:
                xh = mask_s * y + (1 - mask_s) * xh
            return (xh - x) / ti ** 2
        xh = edm.denoiser(x, net, ti.reshape(1, 1))
        if project is not None:
            xh = project(xh)
        elif y is not None:
            xh = mask_s * y + (1 - mask_s) * xh
        elif hpf is not None:
            xh = hpf(xh)
        return (xh - x) / ti ** 2

    for i in range(nb_steps):
        if gamma[i] == 0:
            t_hat = t[i]
        else:
            t_hat = t[i] + gamma[i] * t[i]
            x = x + ((t_hat ** 2 - t[i] ** 2) ** (1 / 2)) * (next(noises) * edm.Snoise)
        d = -t_hat * score(x, t_hat)
        h = t[i + 1] - t_hat
        if t[i + 1] != 0 and order == 2:
            d2 = -t[i + 1] * score(x + h * d, t[i + 1])
            x = x + h * ((1 / 2) * d + (1 / 2) * d2)
        else:
            x = x + h * d
    return x
