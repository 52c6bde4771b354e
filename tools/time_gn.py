"""Times the conv_mode 2 normalise + GELU + operand-layout pass on the six level shapes (B=8): python tools/time_gn.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, aid_b200
from aid_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
for B, Cn, Fd, T in [(8, 64, 64, 4096), (8, 96, 128, 2048), (8, 96, 192, 1024), (8, 128, 256, 512), (8, 128, 320, 256), (8, 256, 384, 128), (8, 256, 448, 64)]:
    x = torch.randn(B, Cn, Fd, T, device=dev)
    ms = C.c_float()
    pf = 0 if T % 128 == 0 else 8
    _lib.check(L.aid_debug_time_gn_tc2(_lib.ptr(x), B, Cn, Fd, T, pf, 10, C.byref(ms)))
    gb = B * Cn * Fd * T * 6 / 1e9
    print(f"bps {os.environ.get('AID_GN_BPS', '16')} B{B} C{Cn} F{Fd} T{T}: {ms.value:7.3f} ms  {gb / ms.value * 1e3:7.0f} GB/s (4 B read + 2 B written per element)", flush=True)
