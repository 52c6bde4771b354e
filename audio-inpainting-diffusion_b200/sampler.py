"""Host-side mirror of the reference's `testing.edm_sampler_inpainting.Sampler` (sampler.py:8-364).

Same constructor and method surface as the reference (consumed by tester_inpainting.py:167,222,415,533):
`predict_inpainting`, `predict_unconditional`, `predict_resample`, `predict`, `apply_mask`,
`prepare_smooth_mask`, plus the externally mutated attributes `.xi`, `.nb_steps`, `.order`.  Select with
    tester.sampler_callable: "audio-inpainting-diffusion_b200.sampler.Sampler"

Differences that are deliberate and documented:
  * The 35-step loop keeps the schedule on the host (same fp32 scalar arithmetic as the reference), so the two
    device->host syncs per step of sampler.py:204,236 disappear.
  * On CUDA tensors each step's element-wise work is one fused kernel (aid_edm_step) and the denoiser call is the
    fused EDM-preconditioned forward; on CPU tensors (only reachable with a non-product denoiser, e.g. the
    host-logic tests) the same updates are written with torch ops in the reference's order.
  * Reconstruction guidance (xi > 0, sampler.py:55-113) needs the denoiser's vector-Jacobian product.  The branch is
    mirrored for any denoiser torch can differentiate (host logic, torch ops); the CUDA denoiser of this package is
    forward-only, so with it xi > 0 raises NotImplementedError -- use xi = 0 (replacement method) there.  The reference
    itself fails for batch > 1 (autograd.grad of a vector norm, sampler.py:78): here the per-clip norms are summed and the
    step size is normalised per clip, which is the reference's arithmetic at batch 1.
  * `prepare_smooth_mask` is vectorised (the reference loops over L samples in Python, sampler.py:311-324).
  * Spectrogram inpainting (sampler.py:271-290, 348-364): on CUDA tensors the STFT -> mask -> inverse STFT degradation and the
    projection `y + x - S(x)` are hand-written kernels (csrc/stft.cu, aid_spectral_mask); on CPU tensors (host-logic tests
    only) the reference's torch.stft / torch.istft calls are used.
"""
import torch

from . import _lib
from .config import cfg_get


class Sampler:
    def __init__(self, model, diff_params, args, rid=False):
        self.model = model
        self.diff_params = diff_params
        self.args = args
        if not cfg_get(args, "tester.diff_params.same_as_training"):
            self.update_diff_params()
        self.order = cfg_get(args, "tester.order")
        self.xi = cfg_get(args, "tester.posterior_sampling.xi")
        use = cfg_get(args, "tester.data_consistency.use")
        kind = cfg_get(args, "tester.data_consistency.type")
        self.data_consistency = use and kind == "always"
        self.data_consistency_end = use and kind == "end"
        if self.data_consistency or self.data_consistency_end:
            self.smooth = bool(cfg_get(args, "tester.data_consistency.smooth"))
        self.nb_steps = cfg_get(args, "tester.T")
        self.rid = rid
        self.noise_source = None  # optional iterator of pre-drawn N(0,1) tensors (prior first), for parity tests
        self.y = self.mask = self.degradation = None
        self._smooth_mask = None
        self._spectral = False      # True while the projection is the spectrogram one (predict_spectrogram_inpainting)
        self._frames = None         # scratch of the CUDA STFT, reused across evaluations

    def update_diff_params(self):
        """sampler.py:43-53"""
        for k in ("sigma_min", "sigma_max", "ro", "sigma_data", "Schurn", "Stmin", "Stmax", "Snoise"):
            setattr(self.diff_params, k, cfg_get(self.args, "tester.diff_params." + k))

    # ---- noise (edm.py:94, sampler.py:212: drawn with the CPU generator, then copied) -----------------
    def _randn(self, shape, device):
        if self.noise_source is not None:
            n = next(self.noise_source)
            assert tuple(n.shape) == tuple(shape)
        else:
            n = torch.randn(shape)
        if torch.device(device).type == "cuda":
            n = n.pin_memory().to(device, non_blocking=True)
        return n.to(device)

    # ---- reference-shaped helpers -------------------------------------------------------------------
    def apply_mask(self, x, mask=None):
        """sampler.py:264-269"""
        if mask is None:
            mask = self.mask
        return mask * x

    def _stft_params(self):
        st = "tester.spectrogram_inpainting.stft."
        if cfg_get(self.args, st + "window") != "hann":
            raise NotImplementedError("Only hann window is implemented for now")     # sampler.py:276
        n_fft, hop, win = (int(cfg_get(self.args, st + k)) for k in ("n_fft", "hop_length", "win_length"))
        return n_fft, hop, win

    def apply_spectral_mask(self, x, y=None):
        """sampler.py:271-290: S(x) = crop(istft(mask * stft(zero-pad(x)))) with self.mask [n_fft/2+1, frames].
        With `y` the projection of the spectrogram mode, y + x - S(x) (sampler.py:361), comes out of the same kernel."""
        n_fft, hop, win = self._stft_params()
        L = x.shape[-1]
        if not x.is_cuda:
            window = torch.hann_window(win).to(x.device)
            xp = torch.nn.functional.pad(x, (0, n_fft - L % n_fft), mode="constant", value=0)
            X = torch.stft(xp, n_fft, hop, win, window, return_complex=True) * self.mask.unsqueeze(0)
            s = torch.istft(X, n_fft, hop, win, window, return_complex=False)[..., 0:L]
            return s if y is None else y + x - s
        if win != n_fft:
            raise NotImplementedError("the CUDA STFT needs win_length == n_fft (the reference's configuration: 1024 / 1024)")
        shape = x.shape
        x2 = x.reshape(-1, L).contiguous().float()
        B = x2.shape[0]
        n_frames = 1 + (L + n_fft - L % n_fft) // hop
        mask = self.mask.to(x.device, torch.float32).contiguous()
        if tuple(mask.shape) != (n_fft // 2 + 1, n_frames):
            raise ValueError(f"spectral mask has shape {tuple(mask.shape)}, the STFT of this input has {(n_fft // 2 + 1, n_frames)}")
        need = B * n_frames * n_fft
        if self._frames is None or self._frames.numel() < need or self._frames.device != x.device:
            self._frames = torch.empty(need, device=x.device, dtype=torch.float32)
        y2 = None if y is None else y.reshape(-1, L).contiguous().float()
        out = torch.empty_like(x2)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().aid_spectral_mask(_lib.ptr(x2), _lib.ptr(y2), _lib.ptr(mask), B, L, n_fft, hop, n_frames,
                                                    _lib.ptr(self._frames), self._frames.numel() * 4, _lib.ptr(out),
                                                    torch.cuda.current_stream(x.device).cuda_stream))
        return out.reshape(shape)

    def prepare_smooth_mask(self, mask, size=10):
        """sampler.py:302-325: half-Hann ramps of `size` samples around every gap of mask[0], broadcast to B rows."""
        hann = torch.hann_window(size * 2)
        hann_left, hann_right = hann[0:size], hann[size::]
        B, N = mask.shape
        m = mask[0].detach().to("cpu")
        new_mask = m.clone()
        prev = torch.cat((torch.ones(1, dtype=m.dtype), m[:-1]))
        for i in torch.nonzero(m != prev)[:, 0].tolist():
            if m[i] == 0:
                new_mask[i - size:i] = hann_right
            if m[i] == 1:
                new_mask[i:i + size] = hann_left
        return new_mask.to(mask.device).unsqueeze(0).expand(B, -1)

    # ---- scores ---------------------------------------------------------------------------------------
    def get_score(self, x, y, t_i, degradation):
        """sampler.py:115-153 (generic torch form; `predict` uses the fused form of the same arithmetic)."""
        if y is None:
            assert degradation is None
            with torch.no_grad():
                x_hat = self.diff_params.denoiser(x, self.model, t_i.unsqueeze(-1))
                if cfg_get(self.args, "tester.filter_out_cqt_DC_Nyq"):
                    x_hat = self.model.CQTransform.apply_hpf_DC(x_hat)
                return (x_hat - x) / t_i ** 2
        if self.xi > 0:
            return self.get_score_rec_guidance(x, y, t_i, degradation)
        with torch.no_grad():
            x_hat = self.diff_params.denoiser(x, self.model, t_i.unsqueeze(-1))
            x_hat = self.proj_convex_set(x_hat.detach())
            return (x_hat.detach() - x) / t_i ** 2

    def _refuse_guidance_on_forward_only_model(self):
        if hasattr(self.model, "denoise_fused"):
            raise NotImplementedError(
                "reconstruction guidance (xi > 0) needs the denoiser's VJP, which this forward-only CUDA denoiser does not "
                "provide; set tester.posterior_sampling.xi = 0 (replacement method)")

    def get_score_rec_guidance(self, x, y, t_i, degradation):
        """sampler.py:55-113: denoise with autograd on, measure ||y - degradation(x_hat)||, step x_hat against its gradient
        w.r.t. x with size t_i * xi / (rms of the gradient), then the optional projection."""
        self._refuse_guidance_on_forward_only_model()
        x = x.detach().requires_grad_()
        with torch.enable_grad():
            x_hat = self.diff_params.denoiser(x, self.model, t_i.unsqueeze(-1))
            if cfg_get(self.args, "tester.filter_out_cqt_DC_Nyq"):
                x_hat = self.model.CQTransform.apply_hpf_DC(x_hat)
            den_rec = degradation(x_hat)
            dim = (1, 2) if y.dim() == 3 else 1
            kind = cfg_get(self.args, "tester.posterior_sampling.norm")
            if kind == "smoothl1":
                norm = torch.nn.functional.smooth_l1_loss(y, den_rec, reduction="sum",
                                                          beta=cfg_get(self.args, "tester.posterior_sampling.smoothl1_beta"))
            else:
                norm = torch.linalg.norm(y - den_rec, dim=dim, ord=kind)
            rec_grads = torch.autograd.grad(outputs=norm.sum(), inputs=x)[0]     # clips are independent: per-clip gradients
        audio_len = cfg_get(self.args, "exp.audio_len")
        gdim = tuple(range(1, rec_grads.dim()))
        normguide = torch.linalg.norm(rec_grads.reshape(rec_grads.shape[0], -1), dim=1).reshape((-1,) + (1,) * len(gdim)) / audio_len ** 0.5
        s = t_i * self.xi / (normguide + 1e-6)
        x_hat = x_hat.detach()
        x_hat_old = x_hat.clone() if self.rid else None
        x_hat = x_hat - s * rec_grads
        x_hat_old_2 = x_hat.clone() if self.rid else None
        if self.data_consistency:
            x_hat = self.proj_convex_set(x_hat.detach())
        score = (x_hat.detach() - x.detach()) / t_i ** 2
        if self.rid:
            return score, x_hat_old, s * rec_grads, x_hat_old_2, x_hat
        return score

    # ---- entry points ---------------------------------------------------------------------------------
    def predict_unconditional(self, shape, device):
        """sampler.py:155-162"""
        self.y = None
        self.degradation = None
        self._smooth_mask = None
        self._spectral = False
        return self.predict(shape, device)

    def predict_resample(self, y, shape, degradation):
        """sampler.py:164-173"""
        self.degradation = degradation
        self.y = y
        self._spectral = False
        return self.predict(shape, y.device)

    def predict_inpainting(self, y_masked, mask):
        """sampler.py:327-346"""
        self.mask = mask.to(y_masked.device)
        self.y = y_masked
        self.degradation = lambda x: self.apply_mask(x)
        self._smooth_mask = None
        self._spectral = False
        if self.data_consistency or self.data_consistency_end:
            if self.smooth:
                smooth_mask = self.prepare_smooth_mask(mask, cfg_get(self.args, "tester.data_consistency.hann_size"))
            else:
                smooth_mask = mask
            smooth_mask = smooth_mask.to(y_masked.device)
            self._smooth_mask, self._proj_y = smooth_mask, y_masked
            self.proj_convex_set = lambda x: smooth_mask * y_masked + (1 - smooth_mask) * x
        return self.predict(self.y.shape, self.y.device)

    def predict_spectrogram_inpainting(self, y_masked, mask):
        """sampler.py:348-364: mask is the real [n_fft/2+1, frames] spectrogram mask, y_masked = apply_spectral_mask(y)."""
        self.mask = mask.to(y_masked.device)
        self.y = y_masked
        self.degradation = lambda x: self.apply_spectral_mask(x)
        self._smooth_mask = None
        self._spectral = False
        if self.data_consistency or self.data_consistency_end:
            self._spectral = True
            self.proj_convex_set = lambda x: self.apply_spectral_mask(x, self.y)     # y + x - S(x)
        return self.predict(self.y.shape, self.y.device)

    # ---- the hot loop ---------------------------------------------------------------------------------
    def predict(self, shape, device):
        """sampler.py:178-262: stochastic 2nd-order EDM sampler; 69 denoiser evaluations for T = 35."""
        conditional = self.y is not None
        if conditional and self.xi > 0:
            return self._predict_guided(shape, device)
        if self.rid:
            raise NotImplementedError("rid=True logging is only defined on the guidance branch of the reference")
        dp = self.diff_params
        shape = tuple(shape)
        dev = torch.device(device)
        t = dp.create_schedule(self.nb_steps)        # host, fp32
        gamma = dp.get_gamma(t)
        x = self._randn(shape, dev) * t[0].to(dev)
        use_proj = conditional and self.data_consistency
        if conditional and not use_proj and not hasattr(self, "proj_convex_set"):
            raise AttributeError("'Sampler' object has no attribute 'proj_convex_set'")  # sampler.py:145 with consistency off
        hpf = (not conditional) and bool(cfg_get(self.args, "tester.filter_out_cqt_DC_Nyq"))
        fused = dev.type == "cuda" and hasattr(self.model, "denoise_fused") and (not conditional or self._smooth_mask is not None or self._spectral)
        ops = _CudaOps(self, dev) if fused else _TorchOps(self)

        for i in range(self.nb_steps):
            if gamma[i] == 0:
                t_hat = t[i]
            else:
                t_hat = t[i] + gamma[i] * t[i]
                eps = self._randn(shape, dev)
                x = ops.add_noise(x, eps, ((t_hat ** 2 - t[i] ** 2) ** (1 / 2)), dp.Snoise)
                del eps
            h = t[i + 1] - t_hat
            second = bool(t[i + 1] != 0) and self.order == 2
            xh = ops.denoise(x, t_hat, hpf)
            d, x_next = ops.step(x, xh, t_hat, h, 0, None, None, conditional)
            if second:
                xh2 = ops.denoise(x_next, t[i + 1], hpf)
                _, x = ops.step(x_next, xh2, t[i + 1], h, 1, d, x, conditional)
            else:
                x = x_next
        if self.data_consistency_end:
            x = self.proj_convex_set(x)
        return x.detach()


    def _predict_guided(self, shape, device):
        """sampler.py:178-262 on the guidance branch (xi > 0): the reference's loop with torch ops, scores from get_score."""
        self._refuse_guidance_on_forward_only_model()
        dp = self.diff_params
        shape = tuple(shape)
        dev = torch.device(device)
        n = self.nb_steps
        if self.rid:
            rid_xt, rid_grads, rid_denoised, rid_grad_update, rid_pocs, rid_xt2 = (torch.zeros((n,) + shape[:2]) for _ in range(6))
        t = dp.create_schedule(n).to(dev)
        x = self._randn(shape, dev) * t[0]
        gamma = dp.get_gamma(t).to(dev)
        for i in range(n):
            if gamma[i] == 0:
                t_hat = t[i]
            else:
                t_hat = t[i] + gamma[i] * t[i]
                x = x + ((t_hat ** 2 - t[i] ** 2) ** (1 / 2)) * (self._randn(shape, dev) * dp.Snoise)
            if self.rid:
                rid_xt[i] = x
            score = self.get_score(x, self.y, t_hat, self.degradation)
            if self.rid:
                score, rid_denoised[i], rid_grads[i], rid_grad_update[i], rid_pocs[i] = score
            d = -t_hat * score
            h = t[i + 1] - t_hat
            if t[i + 1] != 0 and self.order == 2:
                score = self.get_score(x + h * d, self.y, t[i + 1], self.degradation)
                if self.rid:
                    score = score[0]
                x = x + h * ((1 / 2) * d + (1 / 2) * (-t[i + 1] * score))
            else:
                x = x + h * d
            if self.rid:
                rid_xt2[i] = x
        if self.data_consistency_end:
            x = self.proj_convex_set(x)
        if self.rid:
            return (x.detach(), rid_denoised.detach(), rid_grads.detach(), rid_grad_update.detach(), rid_pocs.detach(),
                    rid_xt.detach(), rid_xt2.detach(), t.detach())
        return x.detach()


class _TorchOps:
    """Reference arithmetic with torch ops (sampler.py:214, 141-147, 230-251)."""

    def __init__(self, s):
        self.s = s

    def add_noise(self, x, eps, scale, snoise):
        return x + scale.to(x.device) * (eps * snoise)

    def denoise(self, x, t_i, hpf):
        s = self.s
        with torch.no_grad():
            xh = s.diff_params.denoiser(x, s.model, t_i.to(x.device).reshape(1).unsqueeze(-1))
            if hpf:
                xh = s.model.CQTransform.apply_hpf_DC(xh)
        return xh

    def step(self, xin, xhat, sigma, h, mode, d_prev, xbase, conditional):
        s = self.s
        sigma, h = sigma.to(xin.device), h.to(xin.device)
        if conditional:
            xhat = s.proj_convex_set(xhat.detach())
        score = (xhat.detach() - xin) / sigma ** 2
        d = -sigma * score
        if mode == 0:
            return d, xin + h * d
        return d, xbase + h * ((1 / 2) * d_prev + (1 / 2) * d)


class _CudaOps:
    """Same updates through the C ABI: fused preconditioned denoiser + one element-wise kernel per evaluation."""

    def __init__(self, s, dev):
        self.s, self.dev = s, dev
        self.L = _lib.lib()
        self.mask = None
        self.y = None
        if s._smooth_mask is not None and s.y is not None:  # the projection closes over the y it was built with
            m = s._smooth_mask
            m = m[0] if (m.dim() == 2 and (m.stride(0) == 0 or m.shape[0] == 1)) else m
            self.mask = m.to(dev, torch.float32).contiguous()
            self.y = s._proj_y.to(dev, torch.float32).contiguous()

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def add_noise(self, x, eps, scale, snoise):
        x = x.contiguous()
        with torch.cuda.device(self.dev):
            _lib.check(self.L.aid_edm_add_noise(_lib.ptr(x), _lib.ptr(eps.contiguous()), float(scale) * float(snoise), x.numel(), self._stream()))
        return x

    def denoise(self, x, t_i, hpf):
        s, dp = self.s, self.s.diff_params
        sig = t_i.reshape(1, 1)  # host fp32 scalar: same preconditioning arithmetic as edm.py:143-146
        xh = s.model.denoise_fused(x, dp.cnoise(sig), float(dp.cin(sig)), float(dp.cout(sig)), float(dp.cskip(sig)))
        if hpf:
            xh = s.model.CQTransform.apply_hpf_DC(xh)
        return xh

    def step(self, xin, xhat, sigma, h, mode, d_prev, xbase, conditional):
        xin, xhat = xin.contiguous(), xhat.contiguous()
        x_out = torch.empty_like(xin)
        d_out = torch.empty_like(xin) if mode == 0 else None
        mask = self.mask if conditional else None
        y = self.y if conditional else None
        if conditional and self.s._spectral:      # xhat <- y + xhat - S(xhat) (stft.cu), then the plain step
            xhat, mask, y = self.s.apply_spectral_mask(xhat, self.s.y), None, None
        with torch.cuda.device(self.dev):
            _lib.check(self.L.aid_edm_step(_lib.ptr(xin), _lib.ptr(xhat), _lib.ptr(y), _lib.ptr(mask),
                                           0 if mask is None else mask.numel(), xin.numel(), float(sigma), float(h), mode,
                                           _lib.ptr(d_prev), _lib.ptr(xbase), _lib.ptr(d_out), _lib.ptr(x_out), self._stream()))
        return d_out, x_out
