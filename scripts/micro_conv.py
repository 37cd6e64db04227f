"""Micro-benchmark single convs through irsde_conv2d(engine=1) with IRSDE_TC_DEBUG=1 role-wait counters."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irsde_b200
L = irsde_b200._lib
ctx = L.Context(3, 3, 8, 2, 0, 0, 0)
dev = torch.device("cuda:0")
cases = [(8, 64, 256, 256, 64, 3, 1), (8, 192, 256, 256, 128, 3, 1), (8, 128, 256, 256, 128, 3, 1), (8, 128, 128, 128, 128, 3, 1),
         (8, 64, 256, 256, 384, 1, 0), (8, 1024, 32, 32, 1024, 3, 1), (8, 384, 128, 128, 256, 3, 1)]
for (B, Cin, H, W, Cout, K, pad) in cases:
    x = torch.randn(B, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, K, K, device=dev) * 0.05
    y = torch.empty(B, Cout, H, W, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    print("case", (B, Cin, H, W, Cout, K), flush=True)
    for i in range(2):
        rc = ctx.L.irsde_conv2d(ctx.h, 1, p(x), p(w), None, p(y), B, Cin, H, W, Cout, K, K, 1, pad, 0, 1, None)
        L.check(rc, ctx.h)
