"""Golden vectors at the BENCH shape: the reference's own unet.py (paper network, 186.3 M parameters) at L = 262144.

Run in the build container only:   python tests/golden/make_golden_paper262144.py        (~2 min of CPU)
`EDM.denoiser(x_t, net, sigma)` of the unmodified reference modules (diff_params/edm.py:133-148 ->
networks/unet_cqt_oct_with_projattention_adaLN_2.py:730-845) on one 262144-sample clip at sigma = 1.0 and 0.05, weights =
aid_b200.random_state_dict(paper_22k(262144), seed 1234) loaded with the reference's load_state_dict(strict).  As in
make_golden.py the un-vendored `cqt_nsgt_pytorch` is supplied by oracle/cqt_oracle.py.  Inputs are regenerated from seeds by the
tests (seeded((1, 262144), 0) * sigma); only the outputs are stored (1 MB each, fp32).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets sys.path, installs the CQT restatement, imports the reference modules)

SIGMAS = (1.0, 0.05)


def main():
    aid = mg.aid_b200
    cfg = aid.paper_22k(262144)
    sd = aid.random_state_dict(cfg, seed=1234)
    net = mg.ref_net(cfg, sd)
    args = mg.full_args(cfg)
    edm = mg.RefEDM(args)
    mg.RefSampler(net, edm, args)       # applies tester.diff_params (sampler.py:43-53)
    x = mg.seeded((1, 262144), 0)
    out = {}
    with torch.no_grad():
        for i, sg in enumerate(SIGMAS):
            out[f"paper262144_denoise_{i}"] = edm.denoiser(x * sg, net, torch.tensor([sg])).numpy().astype(np.float32)
            print(i, sg, float(np.abs(out[f"paper262144_denoise_{i}"]).max()), flush=True)
    np.savez(os.path.join(HERE, "golden_paper262144.npz"), **out)


if __name__ == "__main__":
    main()
