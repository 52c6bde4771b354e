"""One fused / two-kernel dilated layer at the level-0 shape (for ncu): python tools/time_comb_one.py [fused]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, aid_b200
from aid_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
fused = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B, Cn, Fd, T, dil = 8, 64, 64, 4096, 2
x = torch.randn(B, Cn, Fd, T, device=dev); w = torch.randn(Cn, Cn, 5, 3, device=dev) * 0.03
gamma = torch.ones(Cn, device=dev); aff = torch.zeros(Cn, device=dev); gate = torch.randn(Cn, device=dev)
out = torch.empty_like(x); st = torch.zeros(B, 8, 2, dtype=torch.float64, device=dev)
ms = C.c_float()
_lib.check(L.aid_debug_dilated_layer(_lib.ptr(x), _lib.ptr(w), B, Cn, Fd, T, dil, _lib.ptr(gamma), _lib.ptr(aff), _lib.ptr(gate), 0.7071, fused, _lib.ptr(out), _lib.ptr(st), C.byref(ms)))
print(ms.value)
