"""Times single convolution launches (device time of the kernel alone) for tuning: python tools/time_conv.py [mode]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, aid_b200
from aid_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
shapes = [  # B, C, F, T, dil   (levels of the paper network at B=8)
    (8, 64, 64, 4096, 2), (8, 96, 128, 2048, 4), (8, 128, 256, 512, 16), (8, 128, 320, 256, 32), (8, 256, 384, 128, 64), (8, 256, 448, 64, 8)]
for B, Cn, Fd, T, dil in shapes:
    a = torch.randn(B, Cn, Fd, T, device=dev); w = torch.randn(Cn, Cn, 5, 3, device=dev) * 0.03
    g = torch.randn(Cn, device=dev); R = torch.randn(B, Cn, Fd, T, device=dev); out = torch.empty_like(R)
    st = torch.zeros(B * 16, dtype=torch.float64, device=dev)
    ms = C.c_float()
    _lib.check(L.aid_debug_time_conv2d(_lib.ptr(a), _lib.ptr(w), B, Cn, Cn, Fd, T, 5, 3, dil, _lib.ptr(g), _lib.ptr(R), 0.7071, _lib.ptr(out), _lib.ptr(st), mode, C.byref(ms)))
    fl = 2.0 * Cn * Cn * 15 * B * Fd * T
    print(f"mode {mode} dbg {os.environ.get('AID_TC_DEBUG','0')} B{B} C{Cn} F{Fd} T{T} d{dil}: {ms.value:8.3f} ms  {fl/ms.value/1e9:8.1f} TFLOP/s (algorithmic)", flush=True)
