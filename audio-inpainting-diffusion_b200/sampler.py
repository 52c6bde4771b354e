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
  * Reconstruction guidance (xi > 0, sampler.py:55-113) needs the denoiser's vector-Jacobian product.  The branch is written
    with torch ops around the denoiser call; this package's CUDA denoiser is differentiable with respect to its input (the
    library's own VJP behind a torch.autograd.Function, unet.py), so `torch.autograd.grad(norm, x)` works exactly as in the
    reference.  The reference itself fails for batch > 1 (autograd.grad of a vector norm, sampler.py:78): here the per-clip
    norms are summed and the step size is normalised per clip, which is the reference's arithmetic at batch 1.
  * `prepare_smooth_mask` is vectorised (the reference loops over L samples in Python, sampler.py:311-324).
  * Noise.  By default it is drawn with torch's CPU generator and copied, exactly as the reference does (edm.py:94,
    sampler.py:212), so seeded runs reproduce the reference's trajectories.  `sampler.device_noise = DeviceNoise(seed, ...)`
    switches to device-resident Philox4x32-10 normals keyed by (seed, stream id, global clip index, draw) (aid_philox_normal,
    restated in oracle/philox_oracle.py): no host work and no H2D copy per step, and a clip's noise does not depend on the batch or
    the rank it is sampled in.  With device noise on a CUDA denoiser of this package the step loop is replayed from two captured
    CUDA graphs (one Heun step, one final Euler step; `use_cuda_graph`, default on): every per-step scalar (noise scale, draw,
    c_in / c_out / c_skip / c_noise, sigma, h) lives in a device table that the graph's head node indexes (aid_sched_select), so
    one graph serves all 35 steps.
  * Spectrogram inpainting (sampler.py:271-290, 348-364): on CUDA tensors the STFT -> mask -> inverse STFT degradation and the
    projection `y + x - S(x)` are hand-written kernels (csrc/stft.cu, aid_spectral_mask); on CPU tensors (host-logic tests
    only) the reference's torch.stft / torch.istft calls are used.
"""
import numpy as np
import torch

from . import _lib
from .config import cfg_get


class DeviceNoise:
    """Device-resident noise stream: seed (64 bit), stream_id (e.g. a per-call counter), clip0 = global index of the batch's first clip."""

    def __init__(self, seed, stream_id=0, clip0=0):
        self.seed, self.stream_id, self.clip0 = int(seed) & (2 ** 64 - 1), int(stream_id) & 0xFFFFFFFF, int(clip0) & 0xFFFFFFFF


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
        self.device_noise = None  # DeviceNoise: Philox normals generated on the GPU instead of torch.randn on the host
        self.use_cuda_graph = True  # with device noise on the fused CUDA path: replay the step loop from captured CUDA graphs
        self._graphs = {}
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

    def apply_spectral_mask(self, x, y=None, mask=None):
        """sampler.py:271-290: S(x) = crop(istft(mask * stft(zero-pad(x)))) with self.mask [n_fft/2+1, frames] (or the given one).
        With `y` the projection of the spectrogram mode, y + x - S(x) (sampler.py:361), comes out of the same kernel."""
        n_fft, hop, win = self._stft_params()
        if mask is None:
            mask = self.mask
        L = x.shape[-1]
        if not x.is_cuda or (torch.is_grad_enabled() and x.requires_grad):
            # CPU tensors (host-logic tests), or the guidance branch differentiating through the degradation (sampler.py:68-78):
            # the reference's own torch.stft / torch.istft calls, which autograd understands
            window = torch.hann_window(win).to(x.device)
            xp = torch.nn.functional.pad(x, (0, n_fft - L % n_fft), mode="constant", value=0)
            X = torch.stft(xp, n_fft, hop, win, window, return_complex=True) * mask.to(x.device).unsqueeze(0)
            s = torch.istft(X, n_fft, hop, win, window, return_complex=False)[..., 0:L]
            return s if y is None else y + x - s
        if win != n_fft:
            raise NotImplementedError("the CUDA STFT needs win_length == n_fft (the reference's configuration: 1024 / 1024)")
        shape = x.shape
        x2 = x.reshape(-1, L).contiguous().float()
        B = x2.shape[0]
        n_frames = 1 + (L + n_fft - L % n_fft) // hop
        mask = mask.to(x.device, torch.float32).contiguous()
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
        if hasattr(self.model, "denoise_fused") and not hasattr(self.model, "_forward_tape"):
            raise NotImplementedError(
                "reconstruction guidance (xi > 0) needs the denoiser's VJP, which this forward-only denoiser does not "
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
        if kind == "smoothl1":
            # the reference runs at any batch size here (a scalar loss) and normalises by ONE norm over the whole batch (sampler.py:83)
            normguide = torch.linalg.norm(rec_grads) / audio_len ** 0.5
        else:
            # sampler.py:78 fails for batch > 1 (autograd.grad of a vector); per-clip norms are the batch-1 arithmetic, clip by clip
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
        use_proj = conditional and self.data_consistency
        if conditional and not use_proj and not hasattr(self, "proj_convex_set"):
            raise AttributeError("'Sampler' object has no attribute 'proj_convex_set'")  # sampler.py:145 with consistency off
        hpf = (not conditional) and bool(cfg_get(self.args, "tester.filter_out_cqt_DC_Nyq"))
        fused = dev.type == "cuda" and hasattr(self.model, "denoise_fused") and (not conditional or self._smooth_mask is not None or self._spectral)
        dn = self.device_noise if self.noise_source is None else None
        if dn is not None and not fused:
            raise NotImplementedError("device_noise needs CUDA tensors and this package's CUDA denoiser (with data consistency when conditional)")
        if dn is not None and self.use_cuda_graph and len(shape) == 2 and self.order in (1, 2):
            x = _GraphedLoop.get(self, shape, dev, conditional, hpf).run(t, gamma, dn)
            if self.data_consistency_end:
                x = self.proj_convex_set(x)
            return x.detach()
        ops = _CudaOps(self, dev) if fused else _TorchOps(self)
        draw = 0
        if dn is not None:
            x = ops.philox(torch.empty(shape, device=dev, dtype=torch.float32), dn, draw, float(t[0]), False)
        else:
            x = self._randn(shape, dev) * t[0].to(dev)

        for i in range(self.nb_steps):
            if gamma[i] == 0:
                t_hat = t[i]
            else:
                t_hat = t[i] + gamma[i] * t[i]
                if dn is not None:
                    draw += 1
                    x = ops.philox(x, dn, draw, float((t_hat ** 2 - t[i] ** 2) ** (1 / 2)) * float(dp.Snoise), True)
                else:
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

    def philox(self, x, dn, draw, scale, accumulate):
        """x[c] = (accumulate ? x[c] : 0) + scale * N(seed, stream, clip0 + c, draw)   (edm.py:94, sampler.py:212-214 on the device)"""
        x = x.contiguous()
        flat = x.reshape(x.shape[0], -1)
        with torch.cuda.device(self.dev):
            _lib.check(self.L.aid_philox_normal(_lib.ptr(flat), flat.shape[0], flat.shape[1], dn.seed, dn.stream_id, dn.clip0, draw,
                                                float(scale), 1 if accumulate else 0, None, self._stream()))
        return x

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


class _GraphedLoop:
    """The 35-step loop of Sampler.predict (sampler.py:201-251) as replays of two captured CUDA graphs.

    Graph A (Heun step): select row -> x += scale * Philox -> xh = D(x; t_hat) [-> HPF | projection] -> d, x1 = Euler ->
    xh2 = D(x1; t_next) [...] -> x = x + h (d + d2) / 2, all on the static buffers below, x updated in place.
    Graph B (last step, t_next = 0, or every step when order == 1): the first half, x = x + h d.
    Nothing of the schedule is baked into the graphs: every scalar comes from row `counter` of a device table
    (row = [noise scale, draw, stream id, clip0 | c_in, c_out, c_skip, c_noise, sigma, h of evaluation 1 | the same of evaluation 2])."""
    ROW, E1, E2, MAX_STEPS = 16, 4, 10, 1024

    @staticmethod
    def get(s, shape, dev, conditional, hpf):
        kind = "spectral" if (conditional and s._spectral) else ("inpaint" if conditional else "uncond")
        key = (tuple(shape), str(dev), kind, bool(hpf), id(s.model))
        g = s._graphs.get(key)
        if g is None:
            s._graphs.clear()          # one live set of static buffers per sampler
            g = s._graphs[key] = _GraphedLoop(s, shape, dev, kind, hpf)
        g.bind_inputs()
        return g

    def __init__(self, s, shape, dev, kind, hpf):
        self.s, self.dev, self.kind, self.hpf, self.shape = s, dev, kind, hpf, tuple(shape)
        self.L = _lib.lib()
        B, n = self.shape
        f = lambda: torch.empty(B, n, device=dev, dtype=torch.float32)
        self.x, self.xh, self.d, self.x1, self.xh2 = f(), f(), f(), f(), f()
        self.hp = f() if hpf else None
        self.y = f() if kind != "uncond" else None
        self.mask = torch.empty(n, device=dev, dtype=torch.float32) if kind == "inpaint" else None
        self.smask = None              # spectral mode: static copy of the [n_fft/2+1, frames] mask
        self.cur = torch.zeros(self.ROW, device=dev, dtype=torch.float32)
        self.counter = torch.zeros(1, device=dev, dtype=torch.int32)
        self.table = torch.zeros(self.MAX_STEPS, self.ROW, device=dev, dtype=torch.float32)
        self.seed = None
        self.graphs = {}               # heun (bool) -> torch.cuda.CUDAGraph
        m = s.model
        m._ensure_weights(dev)
        self.ws = m._workspace(B, dev)
        if hpf:
            nb = _lib.C.c_size_t()
            _lib.check(self.L.aid_cqt_workspace_bytes(m._handle, B, _lib.C.byref(nb)), m._handle)
            self.hws = torch.empty(nb.value, dtype=torch.uint8, device=dev)

    def bind_inputs(self):
        """Copy the current call's y / mask into the static buffers the graphs read."""
        s = self.s
        if self.kind == "inpaint":
            m = s._smooth_mask
            m = m[0] if (m.dim() == 2 and (m.stride(0) == 0 or m.shape[0] == 1)) else m
            if m.dim() != 1:
                raise NotImplementedError("the graphed loop takes one mask shared by the batch (sampler.py:307 uses mask[0] only)")
            self.mask.copy_(m.to(self.dev, torch.float32))
            self.y.copy_(s._proj_y.to(self.dev, torch.float32))
        elif self.kind == "spectral":
            self.y.copy_(s.y.to(self.dev, torch.float32))
            m = s.mask.to(self.dev, torch.float32)
            if self.smask is None or self.smask.shape != m.shape:
                self.smask, self.graphs = torch.empty_like(m), {}
            self.smask.copy_(m)

    # ---- the launches of one step on the current stream --------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _denoise(self, xin, e, out):
        s, m, cur, B = self.s, self.s.model, self.cur, self.shape[0]
        _lib.check(self.L.aid_unet_forward_ds(m._handle, _lib.ptr(xin), _lib.ptr(cur[e + 3:]), 1, _lib.ptr(out), B, _lib.ptr(cur[e:]),
                                              _lib.ptr(self.ws), self.ws.numel(), self._stream()), m._handle)
        if self.hpf:
            _lib.check(self.L.aid_hpf_dc(m._handle, _lib.ptr(out), _lib.ptr(self.hp), B, _lib.ptr(self.hws), self.hws.numel(), self._stream()), m._handle)
            out.copy_(self.hp)
        if self.kind == "spectral":          # xhat <- y + xhat - S(xhat)
            out.copy_(s.apply_spectral_mask(out, self.y, self.smask))

    def _edm(self, xin, xh, e, mode, d_prev, xbase, d_out, x_out):
        mk = self.mask if self.kind == "inpaint" else None
        y = self.y if self.kind == "inpaint" else None
        _lib.check(self.L.aid_edm_step_ds(_lib.ptr(xin), _lib.ptr(xh), _lib.ptr(y), _lib.ptr(mk), 0 if mk is None else mk.numel(), xin.numel(),
                                          _lib.ptr(self.cur[e + 4:]), mode, _lib.ptr(d_prev), _lib.ptr(xbase), _lib.ptr(d_out), _lib.ptr(x_out),
                                          self._stream()))

    def _step(self, heun):
        B, n = self.shape
        _lib.check(self.L.aid_sched_select(_lib.ptr(self.table), self.ROW, _lib.ptr(self.counter), _lib.ptr(self.cur), self._stream()))
        _lib.check(self.L.aid_philox_normal(_lib.ptr(self.x), B, n, self.seed, 0, 0, 0, 0.0, 1, _lib.ptr(self.cur), self._stream()))
        self._denoise(self.x, self.E1, self.xh)
        if heun:
            self._edm(self.x, self.xh, self.E1, 0, None, None, self.d, self.x1)
            self._denoise(self.x1, self.E2, self.xh2)
            self._edm(self.x1, self.xh2, self.E2, 1, self.d, self.x, None, self.x)
        else:
            self._edm(self.x, self.xh, self.E1, 0, None, None, None, self.x)

    def _graph(self, heun):
        g = self.graphs.get(heun)
        if g is None:
            # one eager step first (module loading, shared-memory attributes), on a side stream as capture requires
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                self.x.zero_(); self.counter.zero_()
                self._step(heun)
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step(heun)
            self.graphs[heun] = g
        return g

    def _rows(self, t, gamma, dn):
        """Host-built schedule rows, the same fp32 scalar arithmetic as the eager loop (edm.py:97-128, sampler.py:204-236)."""
        s, dp = self.s, self.s.diff_params
        rows, ints, kinds, draw = [], [], [], 0
        for i in range(s.nb_steps):
            r = [0.0] * self.ROW
            if gamma[i] == 0:
                t_hat = t[i]
            else:
                t_hat = t[i] + gamma[i] * t[i]
                draw += 1
                r[0] = float((t_hat ** 2 - t[i] ** 2) ** (1 / 2)) * float(dp.Snoise)
            h = t[i + 1] - t_hat
            heun = bool(t[i + 1] != 0) and s.order == 2
            for e, sg in ((self.E1, t_hat),) + (((self.E2, t[i + 1]),) if heun else ()):
                sig = sg.reshape(1, 1)
                r[e:e + 6] = [float(dp.cin(sig)), float(dp.cout(sig)), float(dp.cskip(sig)), float(dp.cnoise(sig)), float(sg), float(h)]
            rows.append(r)
            ints.append([draw, dn.stream_id, dn.clip0])
            kinds.append(heun)
        tab = torch.tensor(rows, dtype=torch.float32)
        tab[:, 1:4] = torch.from_numpy(np.array(ints, dtype=np.uint32).view(np.float32))     # draw, stream id, clip0 as bit patterns
        return tab, kinds

    def run(self, t, gamma, dn):
        s = self.s
        if s.nb_steps > self.MAX_STEPS:
            raise ValueError(f"the graphed loop holds at most {self.MAX_STEPS} steps")
        tab, kinds = self._rows(t, gamma, dn)
        with torch.cuda.device(self.dev):
            if self.seed != dn.seed:     # the Philox key is a launch parameter of the captured noise node
                self.seed, self.graphs = dn.seed, {}
            graphs = {h: self._graph(h) for h in sorted(set(kinds))}
            self.table[: tab.shape[0]].copy_(tab.to(self.dev))
            self.counter.zero_()
            _lib.check(self.L.aid_philox_normal(_lib.ptr(self.x), self.shape[0], self.shape[1], dn.seed, dn.stream_id, dn.clip0, 0, float(t[0]),
                                                0, None, self._stream()))
            for heun in kinds:
                graphs[heun].replay()
            return self.x.clone()
