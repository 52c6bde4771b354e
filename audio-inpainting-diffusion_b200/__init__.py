"""B200-native denoiser path for diffusion-based audio inpainting.

Drop-in (same plugin surface, dotted-path selectable) for ONE path of eloimoliner/audio-inpainting-diffusion:
the EDM inpainting sampler loop (testing/edm_sampler_inpainting.py) driving forward passes of the CQT-octave
U-Net denoiser (networks/unet_cqt_oct_with_projattention_adaLN_2.py).  All arithmetic runs in hand-written
sm_100a CUDA kernels behind the C ABI of include/aid_b200.h; this package is the thin host side.
"""
from .config import NetConfig, AttrDict, paper_22k, paper_44k, small_test
from .unet import Unet_CQT_oct_with_attention, random_state_dict, schema_from_lib
from .edm import EDM
from .sampler import Sampler, DeviceNoise
from .masks import prepare_mask, prepare_spectral_mask

__all__ = ["NetConfig", "AttrDict", "paper_22k", "paper_44k", "small_test", "DeviceNoise", "Unet_CQT_oct_with_attention", "random_state_dict",
           "schema_from_lib", "EDM", "Sampler", "prepare_mask", "prepare_spectral_mask"]
