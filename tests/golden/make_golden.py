"""Generates the committed golden vectors by running the REFERENCE's own code (read-only /root/reference).

Run in the build container only:   python tests/golden/make_golden.py
What is the reference's and what is ours:
  * networks/unet_cqt_oct_with_projattention_adaLN_2.py, diff_params/edm.py, testing/edm_sampler_inpainting.py are
    imported unmodified from /root/reference and produce every number stored here;
  * the un-vendored `cqt_nsgt_pytorch` they import is supplied by oracle/cqt_oracle.py (parity unpinned there);
  * weights are the deterministic name-keyed random state dict of aid_b200.random_state_dict (test_mode: gates,
    gammas and biases re-randomised, SURVEY.md finding 6), loaded with the reference's load_state_dict(strict).
Inputs are regenerated from seeds by the tests; only outputs are stored.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), "/root/reference"]

import cqt_oracle  # noqa: E402

cqt_oracle.install_as_cqt_nsgt_pytorch()
import aid_b200  # noqa: E402
from util import seeded  # noqa: E402
from networks.unet_cqt_oct_with_projattention_adaLN_2 import Unet_CQT_oct_with_attention as RefNet  # noqa: E402
from diff_params.edm import EDM as RefEDM  # noqa: E402
from testing.edm_sampler_inpainting import Sampler as RefSampler  # noqa: E402

TESTER = {
    "tester": {
        "T": 35, "order": 2, "filter_out_cqt_DC_Nyq": True,
        "posterior_sampling": {"xi": 0, "norm": 2, "smoothl1_beta": 1},
        "data_consistency": {"use": True, "type": "always", "smooth": True, "hann_size": 50},
        "spectrogram_inpainting": {"stft": {"window": "hann", "n_fft": 1024, "hop_length": 256, "win_length": 1024},
                                   "time_mask_length": 2000, "time_start_idx": "None", "min_masked_freq": 300, "max_masked_freq": 2000},
        "diff_params": {"same_as_training": False, "sigma_data": 0.063, "sigma_min": 1e-4, "sigma_max": 1, "P_mean": -1.2,
                        "P_std": 1.2, "ro": 13, "ro_train": 13, "Schurn": 10, "Snoise": 1.0, "Stmin": 0, "Stmax": 50},
    },
    "diff_params": {"sigma_data": 0.063, "sigma_min": 1e-5, "sigma_max": 10, "P_mean": -1.2, "P_std": 1.2, "ro": 13, "ro_train": 10,
                    "Schurn": 5, "Snoise": 1, "Stmin": 0, "Stmax": 50, "aweighting": {"use_aweighting": False, "ntaps": 101}},
}


def full_args(cfg, T=35):
    a = cfg.to_args()
    a.update(aid_b200.AttrDict.wrap(TESTER))
    a["tester"]["T"] = T
    return a


def ref_net(cfg, sd):
    net = RefNet(cfg.to_args(), "cpu")
    net.load_state_dict(sd, strict=True)
    return net


def inpaint_mask(L, gap, B=1):
    """tester_inpainting.py:231-242 (mask_mode 'long', gap centred)."""
    m = torch.ones(B, L)
    start = L // 2 - gap // 2
    m[..., start:start + gap] = 0
    return m


def main():
    out = {}
    # ---- small network, bare forward ----
    cfg = aid_b200.small_test(16384)
    sd = aid_b200.random_state_dict(cfg, seed=1234)
    net = ref_net(cfg, sd)
    x = seeded((2, cfg.audio_len), 0)
    with torch.no_grad():
        for i, cn in enumerate([0.0613, -0.75, -2.3]):
            out[f"small_fwd_{i}"] = net(x, torch.tensor([[cn]])).numpy()
        x3 = seeded((3, cfg.audio_len), 3, 0.3)
        out["small_fwd_persample"] = net(x3, torch.tensor([[0.05], [-0.8], [-1.7]])).numpy()
    # ---- small network, samplers (reference Sampler + reference EDM) ----
    args = full_args(cfg, T=6)
    smp = RefSampler(net, RefEDM(args), args)
    torch.manual_seed(42)
    out["small_sample_uncond_T6"] = smp.predict_unconditional((2, cfg.audio_len), "cpu").numpy()
    y = seeded((2, cfg.audio_len), 7, 0.063)
    mask = inpaint_mask(cfg.audio_len, 1500)
    torch.manual_seed(43)
    out["small_sample_inpaint_T6"] = smp.predict_inpainting(y * mask, mask).numpy()
    args35 = full_args(cfg, T=35)
    edm = RefEDM(args35)
    RefSampler(net, edm, args35)  # applies tester.diff_params to edm (sampler.py:43-53)
    t = edm.create_schedule(35)
    meta = {"schedule_T35": t.tolist(), "gamma_T35": edm.get_gamma(t).tolist()}
    # ---- paper network (186 M parameters), BASELINE config 1 through EDM.denoiser ----
    cfgp = aid_b200.paper_22k(65536)
    sdp = aid_b200.random_state_dict(cfgp, seed=1234)
    netp = ref_net(cfgp, sdp)
    meta["schema_paper"] = [[k, list(v.shape)] for k, v in netp.state_dict().items()]
    xp = seeded((1, 65536), 0)
    argsp = full_args(cfgp)
    edmp = RefEDM(argsp)
    RefSampler(netp, edmp, argsp)
    with torch.no_grad():
        for i, sg in enumerate([1.0, 0.05]):
            out[f"paper_denoise_{i}"] = edmp.denoiser(xp * sg if i else xp, netp, torch.tensor([sg])).numpy()
    np.savez(os.path.join(HERE, "golden.npz"), **{k: v.astype(np.float32) for k, v in out.items()})
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
