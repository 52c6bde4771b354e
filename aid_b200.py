"""Importable alias of the package directory `audio-inpainting-diffusion_b200` (a hyphen is not a Python identifier).

    import aid_b200            # == importlib.import_module("audio-inpainting-diffusion_b200")

Dotted-path strings resolved through importlib (the reference's dnnlib.call_func_by_name, util.py:292-297) can
use either name.
"""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("audio-inpainting-diffusion_b200")
sys.modules[__name__] = _pkg
for _sub in ("config", "unet", "edm", "sampler", "_lib"):
    sys.modules[f"aid_b200.{_sub}"] = importlib.import_module(f"audio-inpainting-diffusion_b200.{_sub}")
