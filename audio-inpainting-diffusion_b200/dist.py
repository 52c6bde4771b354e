"""Batch-parallel sampling across the GPUs of one box.

Every clip is independent through the denoiser (group-norm statistics and attention are per clip) and through
the xi = 0 sampler, so the batch is partitioned into contiguous slices, one process per GPU, with NO
communication inside the 35-step loop.  The only collective is the result gather at the end
(`torch.distributed.all_gather_into_tensor`: NCCL over NVLink on GPUs, gloo in the CPU tests).

Noise is keyed by (seed, call number, global clip index, draw), so the gathered result does not depend on the number of
ranks, and consecutive predict_* calls of one ShardedSampler draw fresh noise (the reference advances the global RNG
between calls).  On CUDA with this package's denoiser the noise is generated on the device (Philox, aid_philox_normal) and the
step loop is replayed from CUDA graphs; elsewhere (the gloo CPU tests) per-clip torch generators are used.  (The reference
draws one `torch.randn(shape)` for the whole batch, sampler.py:212; that single-stream behaviour is what the un-sharded
Sampler reproduces by default.)
"""
import torch
import torch.distributed as dist

from .sampler import DeviceNoise


def shard_bounds(B, rank, world):
    """Contiguous slice [lo, hi) of a batch of B clips owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ClipNoise:
    """Iterator of N(0,1) tensors [hi-lo, L] whose row b depends only on (seed, call number, global clip index b, draw number)."""

    def __init__(self, seed, lo, hi, L, call=0):
        self.gens = [torch.Generator().manual_seed(((int(seed) * 1000003 + int(call)) * 1000003 + b) % (2 ** 63 - 1)) for b in range(lo, hi)]
        self.L = L

    def __iter__(self):
        return self

    def __next__(self):
        if not self.gens:
            return torch.empty(0, self.L)
        return torch.stack([torch.randn(self.L, generator=g) for g in self.gens])


def gather_clips(x_local, B, group=None):
    """All-gather variable-size contiguous shards back into the [B, L] batch (same result on every rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(B, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    L = x_local.shape[1]
    pad = torch.zeros(nmax, L, dtype=x_local.dtype, device=x_local.device)
    pad[: x_local.shape[0]] = x_local
    out = torch.empty(world * nmax, L, dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * nmax: r * nmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


class ShardedSampler:
    """Wraps a Sampler: each rank samples its slice of the batch, one gather at the end."""

    def __init__(self, sampler, seed=0, group=None, device_noise=True):
        self.sampler, self.seed, self.group = sampler, seed, group
        self.device_noise = device_noise      # False: per-clip host generators even on CUDA
        self.calls = 0                        # folded into the noise key: every predict_* call draws a fresh stream

    def _set_noise(self, lo, hi, L, device):
        """Install this call's noise stream on the wrapped sampler (same call number on every rank)."""
        call, self.calls = self.calls, self.calls + 1
        on_gpu = torch.device(device).type == "cuda" and hasattr(self.sampler.model, "denoise_fused")
        if on_gpu and self.device_noise:
            self.sampler.device_noise, self.sampler.noise_source = DeviceNoise(self.seed, stream_id=call, clip0=lo), None
        else:
            self.sampler.device_noise, self.sampler.noise_source = None, ClipNoise(self.seed, lo, hi, L, call)

    def _clear_noise(self):
        self.sampler.noise_source = self.sampler.device_noise = None

    def _rank_world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def predict_inpainting(self, y_masked, mask):
        """y_masked [B, L] (the whole batch, identical on every rank), mask [1|B, L] -> gathered [B, L]."""
        rank, world = self._rank_world()
        B, L = y_masked.shape
        lo, hi = shard_bounds(B, rank, world)
        self._set_noise(lo, hi, L, y_masked.device)
        try:
            m = mask if mask.shape[0] == 1 else mask[lo:hi]
            if hi > lo:
                x = self.sampler.predict_inpainting(y_masked[lo:hi], m)
            else:
                x = y_masked[lo:hi]
        finally:
            self._clear_noise()
        return gather_clips(x, B, self.group)

    def predict_spectrogram_inpainting(self, y_masked, mask):
        """y_masked [B, L] (the whole batch, identical on every rank), mask [n_fft/2+1, frames] shared by all clips."""
        rank, world = self._rank_world()
        B, L = y_masked.shape
        lo, hi = shard_bounds(B, rank, world)
        self._set_noise(lo, hi, L, y_masked.device)
        try:
            x = self.sampler.predict_spectrogram_inpainting(y_masked[lo:hi], mask) if hi > lo else y_masked[lo:hi]
        finally:
            self._clear_noise()
        return gather_clips(x, B, self.group)

    def predict_unconditional(self, shape, device):
        rank, world = self._rank_world()
        B, L = shape
        lo, hi = shard_bounds(B, rank, world)
        self._set_noise(lo, hi, L, device)
        try:
            x = self.sampler.predict_unconditional((hi - lo, L), device) if hi > lo else torch.empty(0, L, device=device)
        finally:
            self._clear_noise()
        return gather_clips(x, B, self.group)
