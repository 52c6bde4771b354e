"""Golden trajectory of the reference's DEFAULT inpainting mode: reconstruction guidance, xi = 0.25 (conf/tester/inpainting_tester.yaml:32).

Run in the build container only:   python tests/golden/make_golden_guided.py
The unmodified reference Sampler (testing/edm_sampler_inpainting.py:55-113, 178-262) + EDM + unet.py on the CPU, batch 1 (the only
batch size its autograd.grad call accepts), small network, 6 Heun steps = 11 denoiser evaluations each followed by a backward pass
through the network; with and without the data-consistency projection.  Weights / inputs / seeds as in make_golden.py.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    aid = mg.aid_b200
    cfg = aid.small_test(16384)
    sd = aid.random_state_dict(cfg, seed=1234)
    net = mg.ref_net(cfg, sd)
    y = mg.seeded((1, cfg.audio_len), 7, 0.063)
    mask = mg.inpaint_mask(cfg.audio_len, 1500)
    out = {}
    for name, consistency in (("small_sample_guided_T6", True), ("small_sample_guided_noproj_T6", False)):
        args = mg.full_args(cfg, T=6)
        args["tester"]["posterior_sampling"]["xi"] = 0.25
        args["tester"]["data_consistency"]["use"] = consistency
        smp = mg.RefSampler(net, mg.RefEDM(args), args)
        torch.manual_seed(44)
        out[name] = smp.predict_inpainting(y * mask, mask).numpy().astype(np.float32)
        print(name, float(np.abs(out[name]).max()), flush=True)
    np.savez(os.path.join(HERE, "golden_guided.npz"), **out)


if __name__ == "__main__":
    main()
