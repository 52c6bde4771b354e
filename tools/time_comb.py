"""Times one dilated residual layer of conv_mode 2 at 64 channels: operand pass + conv_tc2 against the fused conv_comb_kernel.
python tools/time_comb.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, aid_b200
from aid_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
for B, Fd, T, dil, Cn in [(8, 64, 4096, 2, 64), (8, 128, 2048, 4, 64), (32, 64, 4096, 2, 64), (8, 128, 2048, 4, 96), (8, 192, 1024, 8, 96), (8, 256, 512, 16, 96), (32, 128, 2048, 2, 96)]:
    x = torch.randn(B, Cn, Fd, T, device=dev); w = torch.randn(Cn, Cn, 5, 3, device=dev) * 0.03
    gamma = torch.ones(Cn, device=dev); aff = torch.zeros(Cn, device=dev); gate = torch.randn(Cn, device=dev)
    out = torch.empty_like(x); st = torch.zeros(B, 8, 2, dtype=torch.float64, device=dev)
    for fused in (0, 1):
        ms = C.c_float()
        _lib.check(L.aid_debug_dilated_layer(_lib.ptr(x), _lib.ptr(w), B, Cn, Fd, T, dil, _lib.ptr(gamma), _lib.ptr(aff), _lib.ptr(gate), 0.7071, fused,
                                             _lib.ptr(out), _lib.ptr(st), C.byref(ms)))
        el = B * Cn * Fd * T
        print(f"B{B} C{Cn} F{Fd} T{T} d{dil} fused {fused}: {ms.value:7.3f} ms  {2.0 * Cn * Cn * 15 * B * Fd * T / ms.value / 1e9:7.1f} TFLOP/s  {el * 8 / ms.value / 1e6:6.0f} GB/s at 8 B/element", flush=True)
