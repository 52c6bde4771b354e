"""world_size-2 gloo tests of the batch sharding and the result gather (the only collective on the path)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _spectral_inputs(sampler, y):
    from util import spectral_mask_rect
    smask = spectral_mask_rect(y.shape[-1], gap_ms=30)
    sampler.mask = smask
    return smask, sampler.apply_spectral_mask(y)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import aid_b200
    from aid_b200.dist import ShardedSampler, shard_bounds, gather_clips
    from test_host import _FakeNet, _tester_args
    args = _tester_args(aid_b200, T=5)
    B, L = 5, 2048  # odd batch: ranks get 3 and 2 clips
    y = torch.randn(B, L, generator=torch.Generator().manual_seed(1)) * 0.063
    mask = torch.ones(1, L)
    mask[..., 900:1100] = 0
    s = ShardedSampler(aid_b200.Sampler(_FakeNet(), aid_b200.EDM(args), args), seed=7)
    out = s.predict_inpainting(y * mask, mask)
    smask, ym = _spectral_inputs(s.sampler, y)
    out_s = s.predict_spectrogram_inpainting(ym, smask)
    lo, hi = shard_bounds(B, rank, world)
    g = gather_clips(torch.full((hi - lo, 4), float(rank)), B)
    if rank == 0:
        ret["out"], ret["g"], ret["out_s"] = out, g, out_s
    else:
        ret["out1"], ret["out_s1"] = out, out_s
    dist.destroy_process_group()


def test_sharded_sampling_is_independent_of_world_size(aid):
    from aid_b200.dist import ShardedSampler, shard_bounds
    from test_host import _FakeNet, _tester_args
    assert [shard_bounds(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [shard_bounds(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    out2 = ret["out"]
    assert torch.equal(out2, ret["out1"])  # every rank holds the full gathered batch
    assert ret["g"][:, 0].tolist() == [0.0, 0.0, 0.0, 1.0, 1.0]
    # single process, same seed -> same clips
    args = _tester_args(aid, T=5)
    B, L = 5, 2048
    y = torch.randn(B, L, generator=torch.Generator().manual_seed(1)) * 0.063
    mask = torch.ones(1, L)
    mask[..., 900:1100] = 0
    sh = ShardedSampler(aid.Sampler(_FakeNet(), aid.EDM(args), args), seed=7)
    out1 = sh.predict_inpainting(y * mask, mask)
    assert out1.shape == (B, L) and torch.equal(out1, out2)
    # spectrogram mode: shared [513, frames] mask, clips sharded the same way (second call of the sampler, as in the workers)
    assert torch.equal(ret["out_s"], ret["out_s1"])
    smask, ym = _spectral_inputs(sh.sampler, y)
    assert torch.equal(sh.predict_spectrogram_inpainting(ym, smask), ret["out_s"])


def test_consecutive_calls_draw_fresh_noise(aid):
    """The call number is part of the noise key: two calls of one ShardedSampler differ (the reference advances the global RNG
    between calls), two samplers with the same seed agree call by call."""
    from aid_b200.dist import ShardedSampler
    from test_host import _FakeNet, _tester_args
    args = _tester_args(aid, T=4)
    mk = lambda: ShardedSampler(aid.Sampler(_FakeNet(), aid.EDM(args), args), seed=3)
    a, b = mk(), mk()
    a0, a1 = a.predict_unconditional((2, 1024), "cpu"), a.predict_unconditional((2, 1024), "cpu")
    b0, b1 = b.predict_unconditional((2, 1024), "cpu"), b.predict_unconditional((2, 1024), "cpu")
    assert not torch.equal(a0, a1)
    assert torch.equal(a0, b0) and torch.equal(a1, b1)
    assert a.sampler.noise_source is None and a.sampler.device_noise is None
