"""Golden vectors of the spectrogram-inpainting degradation, produced by the reference's own code (build container only).

    python tests/golden/make_golden_spectral.py      ->  tests/golden/golden_spectral.npz

Calls testing.edm_sampler_inpainting.Sampler.apply_spectral_mask (sampler.py:271-290) unbound, on a namespace that carries
only what the method reads (args.tester.spectrogram_inpainting.stft.*, mask).  Cases: a length that is not a multiple of
n_fft, one that is (the reference then pads a whole extra n_fft), and a smaller transform with a random 0/1 mask.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), "/root/reference"]
import aid_b200  # noqa: E402
import cqt_oracle  # noqa: E402
cqt_oracle.install_as_cqt_nsgt_pytorch()
from testing.edm_sampler_inpainting import Sampler as RefSampler  # noqa: E402
from util import SPECTRAL_CASES, spectral_case  # noqa: E402


def main():
    out = {}
    for name in SPECTRAL_CASES:
        x, mask, n_fft, hop = spectral_case(name)
        ns = types.SimpleNamespace(mask=mask, args=aid_b200.AttrDict.wrap(
            {"tester": {"spectrogram_inpainting": {"stft": {"window": "hann", "n_fft": n_fft, "hop_length": hop, "win_length": n_fft}}}}))
        out[name] = RefSampler.apply_spectral_mask(ns, x).numpy()
    np.savez_compressed(os.path.join(HERE, "golden_spectral.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
