"""Host-side mirror of the reference denoiser's plugin surface.

`Unet_CQT_oct_with_attention(args, device)` is a `torch.nn.Module` with the reference's constructor and
`forward(inputs[B,L], sigma[B|1,1]) -> [B,L]` signature (unet.py:587, 730-737), the reference state-dict key
schema, and a `CQTransform` attribute exposing `fwd / bwd / apply_hpf_DC`
(unet.py:620, sampler.py:63,123).  Select it with
    network.callable: "audio-inpainting-diffusion_b200.unet.Unet_CQT_oct_with_attention"
All arithmetic runs in libaid_b200.so (hand-written sm_100a kernels); torch only owns device memory and streams.
The module is differentiable with respect to its INPUT: when grad is enabled and the input requires grad (the reference's
reconstruction-guidance branch, sampler.py:57-113), forward runs the taped CUDA forward and registers a torch.autograd.Function
whose backward is the library's own vector-Jacobian product (aid_unet_forward_tape / aid_unet_backward).  Parameters and sigma
get no gradient (training is out of scope).
"""
import ctypes as C
import math
import zlib

import torch
import torch.nn as nn

from . import _lib
from .config import NetConfig


def schema_from_lib(cfg: NetConfig):
    """[(name, shape)] of the reference state dict for `cfg`, as the C library builds it (host only, no GPU)."""
    L = _lib.lib()
    h = C.c_void_p()
    c = cfg.to_c()
    _lib.check(L.aid_create(C.byref(c), 0, C.byref(h)))
    try:
        out = []
        name, shape, nd = C.c_char_p(), (C.c_int64 * 4)(), C.c_int()
        for i in range(L.aid_num_weights(h)):
            _lib.check(L.aid_weight_info(h, i, C.byref(name), shape, C.byref(nd)), h)
            out.append((name.value.decode(), tuple(int(shape[j]) for j in range(nd.value))))
        return out
    finally:
        L.aid_destroy(h)


_BUFFERS = ("downsamplerT.kernel", "upsamplerT.kernel")
_CUBIC = [-0.01171875, -0.03515625, 0.11328125, 0.43359375, 0.43359375, 0.11328125, -0.03515625, -0.01171875]


def _fan_in(shape):
    f = 1
    for d in shape[1:]:
        f *= d
    return f


def init_tensor(name, shape, gen, test_mode=False):
    """Reference initialisation (unet.py:20-25, 599-600): kaiming_uniform * sqrt(1/3) == U(-1,1)/sqrt(fan_in);
    gates * 1e-7; biases 0; gamma 1; RFF_freq = 16*randn.

    test_mode re-randomises what the default init makes numerically invisible (SURVEY.md finding 6):
    gates get O(1/sqrt(256)) weights and O(1) biases, gammas are spread around 1, biases are non-zero.
    """
    if name in _BUFFERS:
        return torch.tensor(_CUBIC, dtype=torch.float32)
    if name == "embedding.RFF_freq":
        return 16.0 * torch.randn(shape, generator=gen)
    leaf = name.rsplit(".", 1)[-1]
    is_gate = ".gate" in name
    if leaf == "gamma":
        return 1.0 + 0.25 * torch.randn(shape, generator=gen) if test_mode else torch.ones(shape)
    if leaf == "bias":
        if not test_mode:
            return torch.zeros(shape)
        return (0.6 + 0.2 * torch.randn(shape, generator=gen)) if is_gate else 0.2 * torch.randn(shape, generator=gen)
    u = (torch.rand(shape, generator=gen) * 2 - 1) / math.sqrt(_fan_in(shape))
    if is_gate:
        return u if test_mode else u * (1e-7 * math.sqrt(3.0))
    if test_mode and name.endswith("attn_block.qk.weight"):
        return u * 6.0  # O(1) attention logits: with the default scale the softmax is uniform and q/k bugs are invisible
    return u


def random_state_dict(cfg: NetConfig, seed=1234, test_mode=True):
    """Deterministic, name-keyed random weights in the reference schema (same values on any machine)."""
    sd = {}
    for name, shape in schema_from_lib(cfg):
        gen = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        sd[name] = init_tensor(name, shape, gen, test_mode=test_mode)
    return sd


class _Node(nn.Module):
    pass


class _DenoiseFn(torch.autograd.Function):
    """out = out_scale * net(in_scale * x, c_noise) + skip_scale * x with the library's VJP as backward (input gradient only)."""

    @staticmethod
    def forward(ctx, x, net, c_noise, in_scale, out_scale, skip_scale):
        out = net._forward_tape(x.detach(), c_noise, in_scale, out_scale, skip_scale)
        ctx.net, ctx.gen = net, net._tape_gen
        return out

    @staticmethod
    def backward(ctx, g):
        net = ctx.net
        if net._tape_gen != ctx.gen:
            raise RuntimeError("the denoiser's tape was overwritten by a later differentiable forward: one backward per forward "
                               "(the reference's guidance step, sampler.py:59-78, has this shape)")
        return net._backward(g), None, None, None, None, None


class _HpfFn(torch.autograd.Function):
    """CQT_nsgt.apply_hpf_DC: a real, symmetric circular filter (H[k] real), hence self-adjoint."""

    @staticmethod
    def forward(ctx, x, cqt):
        ctx.cqt = cqt
        return cqt._hpf(x.detach())

    @staticmethod
    def backward(ctx, g):
        return ctx.cqt._hpf(g.contiguous()), None


class CQTDevice:
    """`model.CQTransform` (unet.py:620): fwd / bwd / apply_hpf_DC on the GPU through the C ABI."""

    def __init__(self, owner):
        self._o = owner

    def _ws(self, B, dev):
        L = _lib.lib()
        n = C.c_size_t()
        _lib.check(L.aid_cqt_workspace_bytes(self._o._handle, B, C.byref(n)), self._o._handle)
        return torch.empty(n.value, dtype=torch.uint8, device=dev)

    def _layout(self, B):
        L = _lib.lib()
        no = self._o.cfg.num_octs
        offs, frames = (C.c_int64 * (no + 1))(), (C.c_int32 * no)()
        _lib.check(L.aid_cqt_layout(self._o._handle, B, offs, frames), self._o._handle)
        return list(offs), list(frames)

    def fwd(self, x):
        """real [B,1,L] -> list (ascending frequency) of complex64 [B,1,bins,T_o]   (unet.py:743)"""
        o = self._o
        o._ensure_handle(x.device)
        B = x.shape[0]
        assert x.shape[1] == 1 and x.shape[-1] == o.cfg.audio_len
        xin = x.reshape(B, -1).contiguous().float()
        offs, frames = self._layout(B)
        coef = torch.empty(offs[-1], dtype=torch.float32, device=x.device)
        ws = self._ws(B, x.device)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().aid_cqt_fwd(o._handle, _lib.ptr(xin), _lib.ptr(coef), B, _lib.ptr(ws), ws.numel(), st), o._handle)
        out = []
        for i, T in enumerate(frames):
            c = coef[offs[i]:offs[i + 1]].view(B, 2, o.cfg.bins_per_oct, T)
            out.append(torch.complex(c[:, 0], c[:, 1]).unsqueeze(1))
        return out

    def bwd(self, coefs):
        """list of complex [B,1,bins,T_o] -> real [B,1,L]   (unet.py:841)"""
        o = self._o
        dev = coefs[0].device
        o._ensure_handle(dev)
        B = coefs[0].shape[0]
        offs, frames = self._layout(B)
        coef = torch.empty(offs[-1], dtype=torch.float32, device=dev)
        for i, c in enumerate(coefs):
            assert c.shape[-1] == frames[i], "octave frame count mismatch"
            dst = coef[offs[i]:offs[i + 1]].view(B, 2, o.cfg.bins_per_oct, frames[i])
            cc = c.reshape(B, o.cfg.bins_per_oct, frames[i])
            dst[:, 0].copy_(cc.real)
            dst[:, 1].copy_(cc.imag)
        x = torch.empty(B, o.cfg.audio_len, dtype=torch.float32, device=dev)
        ws = self._ws(B, dev)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().aid_cqt_bwd(o._handle, _lib.ptr(coef), _lib.ptr(x), B, _lib.ptr(ws), ws.numel(), st), o._handle)
        return x.unsqueeze(1)

    def apply_hpf_DC(self, x):
        """[B,L'] (L' <= L, zero padded) -> [B,L']   (sampler.py:63,123); differentiable (guidance branch, sampler.py:62-63)"""
        if torch.is_grad_enabled() and x.requires_grad:
            return _HpfFn.apply(x, self)
        return self._hpf(x)

    def _hpf(self, x):
        o = self._o
        o._ensure_handle(x.device)
        Lin, L = x.shape[-1], o.cfg.audio_len
        if Lin > L:
            raise ValueError("Input signal is longer than the maximum length")
        xin = x.reshape(-1, Lin).float()
        if Lin < L:
            xin = torch.nn.functional.pad(xin, (0, L - Lin))
        xin = xin.contiguous()
        B = xin.shape[0]
        out = torch.empty_like(xin)
        ws = self._ws(B, x.device)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().aid_hpf_dc(o._handle, _lib.ptr(xin), _lib.ptr(out), B, _lib.ptr(ws), ws.numel(), st), o._handle)
        return out[:, :Lin].reshape(x.shape)


class Unet_CQT_oct_with_attention(nn.Module):
    """Drop-in for networks.unet_cqt_oct_with_projattention_adaLN_2.Unet_CQT_oct_with_attention."""

    def __init__(self, args, device, conv_mode=None):
        super().__init__()
        self.args = args
        self.cfg = args if isinstance(args, NetConfig) else NetConfig.from_args(args, conv_mode=conv_mode)
        if conv_mode is not None:
            self.cfg.conv_mode = int(conv_mode)
        self.device = torch.device(device)
        self.depth = self.cfg.num_octs
        self.emb_dim = self.cfg.emb_dim
        self.bins_per_oct, self.num_octs = self.cfg.bins_per_oct, self.cfg.num_octs
        self._handle = None
        self._handle_dev = None
        self._weights_loaded = False
        self._ws = {}
        self._tape_ws = None
        self._tape_gen = 0
        gen = torch.Generator().manual_seed(torch.initial_seed() % (2 ** 31))
        for name, shape in schema_from_lib(self.cfg):
            node = self
            *path, leaf = name.split(".")
            for part in path:
                if not hasattr(node, part):
                    node.add_module(part, _Node())
                node = getattr(node, part)
            t = init_tensor(name, shape, gen)
            if name in _BUFFERS:
                node.register_buffer(leaf, t)
            else:
                node.register_parameter(leaf, nn.Parameter(t, requires_grad=(name != "embedding.RFF_freq")))
        self.CQTransform = CQTDevice(self)

    # ---- handle / weights -------------------------------------------------------------------------
    def _dev_index(self, dev):
        dev = torch.device(dev)
        if dev.type != "cuda":
            raise _lib.AidError("this denoiser runs on a CUDA device only (no CPU fallback); got device " + str(dev))
        return dev.index if dev.index is not None else torch.cuda.current_device()

    def _ensure_handle(self, dev):
        idx = self._dev_index(dev)
        if self._handle is not None and self._handle_dev == idx:
            return
        self._release()
        h = C.c_void_p()
        c = self.cfg.to_c()
        _lib.check(_lib.lib().aid_create(C.byref(c), idx, C.byref(h)))
        self._handle, self._handle_dev, self._weights_loaded = h, idx, False

    def _ensure_weights(self, dev):
        self._ensure_handle(dev)
        if self._weights_loaded:
            return
        L = _lib.lib()
        for name, t in self.state_dict().items():
            if name in _BUFFERS:
                continue
            t = t.detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(L.aid_load_weight(self._handle, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()), self._handle)
        _lib.check(L.aid_finalize(self._handle), self._handle)
        self._weights_loaded = True

    def refresh_weights(self):
        """Re-upload the parameters after they were modified in place."""
        self._release()

    def _release(self):
        if self._handle is not None:
            _lib.lib().aid_destroy(self._handle)
        self._handle, self._handle_dev, self._weights_loaded = None, None, False
        self._ws = {}
        self._tape_ws = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def load_state_dict(self, state_dict, strict=True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self._release()
        return r

    def _workspace(self, B, dev):
        key = (B, str(dev))
        if key not in self._ws:
            n = C.c_size_t()
            _lib.check(_lib.lib().aid_workspace_bytes(self._handle, B, C.byref(n)), self._handle)
            self._ws = {key: torch.empty(n.value, dtype=torch.uint8, device=dev)}
        return self._ws[key]

    # ---- forward ----------------------------------------------------------------------------------
    def _prep(self, x, c_noise):
        if x.dim() != 2 or x.shape[1] != self.cfg.audio_len:
            raise AssertionError("bad shapes")  # unet.py:844
        x = x.float().contiguous()
        self._ensure_weights(x.device)
        cn = c_noise.detach().reshape(-1).to(device=x.device, dtype=torch.float32).contiguous()
        if cn.numel() not in (1, x.shape[0]):
            raise ValueError("sigma must have 1 or B entries")
        return x, cn

    def _forward_tape(self, x, c_noise, in_scale, out_scale, skip_scale):
        """Taped forward (aid_unet_forward_tape): the workspace keeps what `_backward` needs until the next taped forward."""
        x, cn = self._prep(x, c_noise)
        B, dev = x.shape[0], x.device
        L = _lib.lib()
        need = C.c_size_t()
        _lib.check(L.aid_vjp_workspace_bytes(self._handle, B, C.byref(need)), self._handle)
        if self._tape_ws is None or self._tape_ws.numel() < need.value or self._tape_ws.device != dev:
            self._tape_ws = None
            self._tape_ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.aid_unet_forward_tape(self._handle, _lib.ptr(x), _lib.ptr(cn), cn.numel(), _lib.ptr(out), B, float(in_scale),
                                               float(out_scale), float(skip_scale), _lib.ptr(self._tape_ws), self._tape_ws.numel(), st),
                       self._handle)
        self._tape_gen += 1
        self._tape_x = x          # keeps the input alive (the backward does not read it, the caller's graph might)
        return out

    def _backward(self, g):
        """grad_x for the last taped forward (aid_unet_backward)."""
        g = g.float().contiguous()
        gx = torch.empty_like(g)
        with torch.cuda.device(g.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().aid_unet_backward(self._handle, _lib.ptr(g), _lib.ptr(gx), st), self._handle)
        return gx

    def denoise_fused(self, x, c_noise, in_scale=1.0, out_scale=1.0, skip_scale=0.0, out=None):
        """out = out_scale * net(in_scale * x, c_noise) + skip_scale * x  (EDM.denoiser fused, edm.py:133-148)."""
        if torch.is_grad_enabled() and x.requires_grad:
            return _DenoiseFn.apply(x, self, c_noise, float(in_scale), float(out_scale), float(skip_scale))
        if x.dim() != 2 or x.shape[1] != self.cfg.audio_len:
            raise AssertionError("bad shapes")  # unet.py:844
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        dev = x.device
        self._ensure_weights(dev)
        B = x.shape[0]
        cn = c_noise.reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
        if cn.numel() not in (1, B):
            raise ValueError("sigma must have 1 or B entries")
        if out is None:
            out = torch.empty_like(x)
        ws = self._workspace(B, dev)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().aid_unet_forward(self._handle, _lib.ptr(x), _lib.ptr(cn), cn.numel(), _lib.ptr(out), B,
                                                   float(in_scale), float(out_scale), float(skip_scale),
                                                   _lib.ptr(ws), ws.numel(), st), self._handle)
        return out

    def saturation_counts(self, enable=None):
        """(activation, weight) operand values clamped to the finite fp16 range in conv_mode 2 (aid_debug_saturation).
        enable=True resets the activation counter and counts during the following forwards, enable=False stops, None only reads."""
        if self._handle is None or not self._weights_loaded:
            raise _lib.AidError("saturation_counts needs an uploaded model (run a forward or _ensure_weights first)")
        a, w = C.c_uint64(), C.c_uint64()
        L = _lib.lib()
        if enable is None:
            _lib.check(L.aid_debug_saturation(self._handle, 1 if getattr(self, "_count_sat", False) else 0, C.byref(a), C.byref(w)), self._handle)
        else:
            self._count_sat = bool(enable)
            _lib.check(L.aid_debug_saturation(self._handle, 1 if enable else 0, C.byref(a), C.byref(w)), self._handle)
        return int(a.value), int(w.value)

    def set_fusion(self, init_blocks=None, dilated_layers=None, out_blocks=None, upsampling=None):
        """Parity tests: run the un-fused twins of the fused conv_mode 2 kernels in later forwards (aid_debug_fusion); None leaves a switch."""
        if self._handle is None:
            raise _lib.AidError("set_fusion needs an uploaded model (run a forward or _ensure_weights first)")
        enc = lambda v: -1 if v is None else int(bool(v))
        _lib.check(_lib.lib().aid_debug_fusion(self._handle, enc(init_blocks), enc(dilated_layers), enc(out_blocks), enc(upsampling)), self._handle)

    def forward_with_probes(self, inputs, sigma):
        """Debug/parity helper: forward plus the per-block intermediates {"enc<i>", "mid", "dec<i>"} (see aid_debug_probe)."""
        cfg, B, dev = self.cfg, inputs.shape[0], inputs.device
        self._ensure_weights(dev)
        no, bins = cfg.num_octs, cfg.bins_per_oct
        _, frames = self.CQTransform._layout(1)
        T = lambda lvl: frames[no - 1 - lvl]
        shapes = {f"enc{i}": (B, cfg.Ns[i], bins * (i + 1), T(i)) for i in range(no)}
        shapes["mid"] = (B, cfg.Ns[no - 1], bins * no, T(no - 1))
        for i in range(no):
            j = no - 1 - i
            shapes[f"dec{i}"] = (B, cfg.Ns[0] if j == 0 else cfg.Ns[j - 1], bins * (j + 1), T(j))
        bufs = {k: torch.empty(s, dtype=torch.float32, device=dev) for k, s in shapes.items()}
        L = _lib.lib()
        try:
            for k, b in bufs.items():
                _lib.check(L.aid_debug_probe(self._handle, k.encode(), _lib.ptr(b)), self._handle)
            out = self.denoise_fused(inputs, sigma)
            torch.cuda.synchronize(dev)
        finally:
            for k in bufs:
                L.aid_debug_probe(self._handle, k.encode(), None)
        return out, bufs

    def forward(self, inputs, sigma):
        """inputs [B,T] time-domain signal, sigma [B,1] or [1,1] noise-level embedding input (c_noise)."""
        return self.denoise_fused(inputs, sigma)      # differentiable with respect to `inputs` (see _DenoiseFn)
