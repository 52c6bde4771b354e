"""Debug helper for the tcgen05 conv kernels: small shapes, error broken down by row / pixel phase / cout phase."""
import math, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import aid_b200
from aid_b200 import _lib as L
from util import rel_l2, seeded
import torch.nn.functional as F
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=200, precision=2, sci_mode=False)
def run(B, Cin, Cout, Fd, T, KF, KT, dil, mode=3):
    a = seeded((B, Cin, Fd, T), 1); w = seeded((Cout, Cin, KF, KT), 2, 1.0 / math.sqrt(Cin * KF * KT))
    ar, wr = (a * 16).half().float() / 16, (w * 1024).half().float() / 1024
    ref = F.conv2d(ar.double(), wr.double(), padding="same", dilation=(dil, 1))
    out = torch.full((B, Cout, Fd, T), float("nan"), device=dev)
    ad, wd = a.to(dev), w.to(dev)
    L.check(L.lib().aid_op_conv2d(L.ptr(ad), L.ptr(wd), B, Cin, Cout, Fd, T, KF, KT, dil, None, None, None, 1.0, 0.0, L.ptr(out), None, mode, None))
    torch.cuda.synchronize()
    o = out.cpu().double()
    r = rel_l2(o, ref)
    print(f"B{B} Ci{Cin} Co{Cout} F{Fd} T{T} k{KF}x{KT} d{dil}: rel {r:.3e}", flush=True)
    if r > 1e-4 and "-v" in sys.argv:
        e = (o - ref).abs()
        print("   err by (b,row):", e.amax(dim=(1, 3)))
        Tq = (T // 8) * 8
        print("   err by (t%8 x cout%8):\n", e[..., :Tq].reshape(B, Cout // 8, 8, Fd, Tq // 8, 8).amax(dim=(0, 1, 3, 4)).t())
        if KF * KT == 1:
            # which input chunks (8 channels) are missing/duplicated: regress out on per-chunk partial sums for b0,row0
            parts = torch.stack([F.conv2d(ar[:, c0:c0 + 8].double(), wr[:, c0:c0 + 8].double()) for c0 in range(0, Cin, 8)])  # [chunks,B,Co,F,T]
            for f in range(min(Fd, 3)):
                A = parts[:, 0, :, f, :].reshape(parts.shape[0], -1).t()   # [Co*T, chunks]
                y = o[0, :, f, :].reshape(-1, 1)
                sol = torch.linalg.lstsq(A, y).solution.flatten()
                print(f"   row {f}: per-chunk coefficients", sol)
import itertools
if "-grid" in sys.argv:
    for (KF, KT), T, Fd in itertools.product([(1, 1), (5, 3)], [64, 128, 256], [1, 2, 8]):
        line = f"k{KF}x{KT} T{T} F{Fd}: "
        for Ci, Co in itertools.product([16, 64, 128, 192], [16, 64, 128, 256]):
            a = seeded((1, Ci, Fd, T), 1); w = seeded((Co, Ci, KF, KT), 2, 1.0 / math.sqrt(Ci * KF * KT))
            ar, wr = (a * 16).half().float() / 16, (w * 1024).half().float() / 1024
            ref = F.conv2d(ar.double(), wr.double(), padding="same", dilation=(1, 1))
            out = torch.full((1, Co, Fd, T), float("nan"), device=dev)
            ad, wd = a.to(dev), w.to(dev)
            L.check(L.lib().aid_op_conv2d(L.ptr(ad), L.ptr(wd), 1, Ci, Co, Fd, T, KF, KT, 1, None, None, None, 1.0, 0.0, L.ptr(out), None, 3, None))
            torch.cuda.synchronize()
            r = rel_l2(out.cpu().double(), ref)
            line += f" {Ci}>{Co}:{'ok' if r < 1e-4 else f'{r:.2f}'}"
        print(line, flush=True)
    sys.exit(0)
def blockmap(B, Cin, Cout, Fd, T, KF, KT, dil):
    a = seeded((B, Cin, Fd, T), 1); w = seeded((Cout, Cin, KF, KT), 2, 1.0 / math.sqrt(Cin * KF * KT))
    ar, wr = (a * 16).half().float() / 16, (w * 1024).half().float() / 1024
    ref = F.conv2d(ar.double(), wr.double(), padding="same", dilation=(dil, 1))
    out = torch.full((B, Cout, Fd, T), float("nan"), device=dev)
    ad, wd = a.to(dev), w.to(dev)
    L.check(L.lib().aid_op_conv2d(L.ptr(ad), L.ptr(wd), B, Cin, Cout, Fd, T, KF, KT, dil, None, None, None, 1.0, 0.0, L.ptr(out), None, 3, None))
    torch.cuda.synchronize()
    o = out.cpu().double()
    print(f"B{B} Ci{Cin} Co{Cout} F{Fd} T{T} k{KF}x{KT}: rel {rel_l2(o, ref):.3e}  |out| {o.norm():.3f} |ref| {ref.norm():.3f} nan {int(torch.isnan(o).sum())}")
    e = (o - ref).abs()[0]   # [Co, F, T]
    cb = 8 if Cout <= 64 else 32
    m = e.reshape(Cout // cb, cb, Fd, T // 32, 32).amax(dim=(1, 4))   # [Co blocks, F, T/32]
    for f in range(Fd):
        print(f"   row {f}: max err by (cout block of {cb}) x (32-pixel quarter):")
        print(m[:, f, :])
    print("   sample out:", o[0, :4, 0, :4].flatten().tolist())
    print("   sample ref:", ref[0, :4, 0, :4].flatten().tolist())
def ident(Cin, Cout, T=128):
    """W = identity-like (w[n, c] = 1 if c == n % Cin), a[c, t] = c + 64 * (t % 16): out[n, t] reveals which channel was picked."""
    a = torch.zeros(1, Cin, 1, T)
    for c in range(Cin):
        a[0, c, 0, :] = c + 64.0 * (torch.arange(T) % 16)
    w = torch.zeros(Cout, Cin, 1, 1)
    for n in range(Cout):
        w[n, n % Cin, 0, 0] = 1.0
    out = torch.full((1, Cout, 1, T), float("nan"), device=dev)
    ad, wd = a.to(dev), w.to(dev)
    L.check(L.lib().aid_op_conv2d(L.ptr(ad), L.ptr(wd), 1, Cin, Cout, 1, T, 1, 1, 1, None, None, None, 1.0, 0.0, L.ptr(out), None, 3, None))
    torch.cuda.synchronize()
    o = out.cpu()[0, :, 0, :]   # [Cout, T]
    print(f"identity test Ci{Cin} Co{Cout}: picked channel (value % 64) for n = 0..{Cout-1} at t = 0..11 (expected n % Cin); -1 = zero")
    for n in range(min(Cout, 24)):
        row = []
        for t in range(12):
            v = float(o[n, t])
            row.append(-1 if v == 0 and not (n % Cin == 0 and t % 16 == 0) else int(round(v)) % 64)
        print(f"   n={n:3d}: {row}   t-code {[int(round(float(o[n, t]))) // 64 for t in range(12)]}")
ident(64, 16); ident(64, 64); ident(16, 16)
