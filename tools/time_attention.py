"""Device time of the attention core on the paper network's four attention shapes at B=8: python tools/time_attention.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, aid_b200
from aid_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
for tc in (0, 1):
    tot = 0.0
    for B, Fd, T, n in [(8, 320, 256, 2), (8, 384, 128, 2), (8, 448, 64, 3)]:
        h = torch.randn(B, 8, Fd, T, device=dev); qk = torch.randn(B, 16 * Fd, T, device=dev) * 0.3; out = torch.empty_like(h)
        for _ in range(3):
            _lib.check(L.aid_op_attention_mode(_lib.ptr(h), _lib.ptr(qk), B, 8, Fd, T, _lib.ptr(out), tc, None))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            _lib.check(L.aid_op_attention_mode(_lib.ptr(h), _lib.ptr(qk), B, 8, Fd, T, _lib.ptr(out), tc, None))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20; tot += n * ms
        print(f"{'tcgen05' if tc else 'simt   '} B{B} F{Fd} T{T}: {ms:.3f} ms per launch ({n} per forward)")
    print(f"{'tcgen05' if tc else 'simt   '}: {tot:.3f} ms per forward at B=8 (7 attention blocks)")
