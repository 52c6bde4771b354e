"""Multi-GPU path on hardware (needs >= 2 GPUs; skipped otherwise): ShardedSampler with the real CUDA denoiser over NCCL --
one process per GPU, batch sharded, no collective inside the step loop, one all-gather of the results."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import aid_b200
    from aid_b200.dist import ShardedSampler
    from test_host import _tester_args
    from util import seeded
    cfg = aid_b200.small_test(16384, conv_mode=2)
    net = aid_b200.Unet_CQT_oct_with_attention(cfg, dev)
    net.load_state_dict(aid_b200.random_state_dict(cfg, seed=1234))
    args = _tester_args(aid_b200, T=4)
    B, L = 5, cfg.audio_len          # odd batch: ranks get 3 and 2 clips
    y = seeded((B, L), 7, 0.063).to(dev)
    mask = torch.ones(1, L, device=dev)
    mask[..., 8000:8600] = 0
    sh = ShardedSampler(aid_b200.Sampler(net, aid_b200.EDM(args), args), seed=11)
    out = sh.predict_inpainting(y * mask, mask)
    outu = sh.predict_unconditional((B, L), dev)
    ret[f"inp{rank}"], ret[f"unc{rank}"] = out.cpu(), outu.cpu()
    dist.destroy_process_group()


def test_sharded_sampler_nccl_two_ranks_equals_one_rank(aid, cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from aid_b200.dist import ShardedSampler
    from test_host import _tester_args
    from util import seeded, rel_l2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 2000, ret), nprocs=2, join=True)
    assert torch.equal(ret["inp0"], ret["inp1"]) and torch.equal(ret["unc0"], ret["unc1"])   # every rank holds the gathered batch
    # one rank, same seed: the same clips (device Philox noise is keyed by the global clip index)
    cfg = aid.small_test(16384, conv_mode=2)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(aid.random_state_dict(cfg, seed=1234))
    args = _tester_args(aid, T=4)
    B, L = 5, cfg.audio_len
    y = seeded((B, L), 7, 0.063).to(cuda)
    mask = torch.ones(1, L, device=cuda)
    mask[..., 8000:8600] = 0
    sh = ShardedSampler(aid.Sampler(net, aid.EDM(args), args), seed=11)
    one = sh.predict_inpainting(y * mask, mask)
    oneu = sh.predict_unconditional((B, L), cuda)
    assert rel_l2(ret["inp0"], one) < 2e-4 and rel_l2(ret["unc0"], oneu) < 2e-4       # run-to-run noise of conv_mode 2 only
