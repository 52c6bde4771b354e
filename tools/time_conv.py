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
def run(B, Ci, Co, Fd, T, K, dil, useR=True):
    KF, KT = (5, 3) if K == 5 else (1, 1)
    a = torch.randn(B, Ci, Fd, T, device=dev)  # mode 4 reads the same buffers as channels-last: timing only
    w = torch.randn(Co, Ci, KF, KT, device=dev) * 0.03
    g = torch.randn(Co, device=dev); R = torch.randn(B, Co, Fd, T, device=dev) if useR else None; out = torch.empty(B, Co, Fd, T, device=dev)
    st = torch.zeros(B * 16, dtype=torch.float64, device=dev) if os.environ.get('NOSTATS') is None else None
    ms = C.c_float()
    _lib.check(L.aid_debug_time_conv2d(_lib.ptr(a), _lib.ptr(w), B, Ci, Co, Fd, T, KF, KT, dil, _lib.ptr(g), _lib.ptr(R), 0.7071, _lib.ptr(out), _lib.ptr(st), mode, C.byref(ms)))
    fl = 2.0 * Ci * Co * KF * KT * B * Fd * T
    gb = 4.0 * B * Fd * T * (Ci + Co * (2 if useR else 1)) / 1e9
    if int(os.environ.get("AID_TC_DEBUG", "0")) & 2048:
        buf = (C.c_uint64 * 16)()
        _lib.check(L.aid_debug_tc2_profile(buf))
        n = max(1, buf[9]) * 2      # two launches (warm-up + timed) were summed
        names = ["mma:wait_tmem_empty", "mma:wait_a_full", "mma:wait_b_full", "mma:issue", "mma:total", "A:wait_empty", "B:wait_empty", "epi:wait_full", "epi:total"]
        print("   profile (k cycles per CTA): " + ", ".join(f"{nm} {buf[i] / buf[9] / 1e3:.1f}" for i, nm in enumerate(names)) + f", mma:tile_setup {buf[10] / buf[9] / 1e3:.1f}")
    print(f"mode {mode} dbg {os.environ.get('AID_TC_DEBUG','0')} nt {os.environ.get('AID_TC_NTILE','0')} nostats {int(os.environ.get('NOSTATS') is not None)} {K}x B{B} Ci{Ci} Co{Co} F{Fd} T{T} d{dil} R{int(useR)}: {ms.value:8.3f} ms  {fl/ms.value/1e9:8.1f} TFLOP/s  {gb/ms.value*1e3:7.0f} GB/s (algorithmic)", flush=True)

which = sys.argv[2] if len(sys.argv) > 2 else "all"
if os.environ.get("TC_SHAPES"):   # "B,C,F,T,dil;B,C,F,T,dil;..."  (5x3 layers)
    for sh in os.environ["TC_SHAPES"].split(";"):
        B, Cn, Fd, T, dil = (int(v) for v in sh.split(","))
        run(B, Cn, Cn, Fd, T, 5, dil, useR=os.environ.get('NOR') is None)
    sys.exit(0)
if which == "thin":   # the 2 -> N pyramid convolutions (5x3) of the seven levels, CUDA-core thin-channel kernel (mode 0)
    for B, Co, Fd, T in [(8, 64, 64, 4096), (8, 96, 128, 2048), (8, 96, 192, 1024), (8, 128, 256, 512), (8, 128, 320, 256), (8, 256, 384, 128), (8, 256, 448, 64)]:
        run(B, 2, Co, Fd, T, 5, 1)
    sys.exit(0)
if which in ("all", "5x3"):
    for B, Cn, Fd, T, dil in shapes:
        run(B, Cn, Cn, Fd, T, 5, dil)
if which in ("all", "1x1"):
    for B, Ci, Co, Fd, T, useR in [(8, 64, 64, 64, 4096, True), (8, 64, 96, 128, 2048, False), (8, 128, 64, 64, 4096, True), (8, 192, 64, 128, 2048, False),
                                   (8, 256, 256, 64, 64, True), (8, 3584, 7168, 1, 64, False), (8, 2560, 5120, 1, 256, False)]:
        run(B, Ci, Co, Fd, T, 1, 1, useR)
