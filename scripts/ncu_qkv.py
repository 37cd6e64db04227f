"""One to_qkv-shaped tcgen05 conv launch (8x128x256x256 -> 384 with the q-softmax epilogue) for an ncu source-level capture."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irsde_b200
L = irsde_b200._lib
ctx = L.Context(3, 3, 8, 2, L.NET_CONDITIONAL, L.PREC_FP32, 0)
dev = torch.device("cuda:0")
B, Cin, H, W, Cout = 8, 128, 256, 256, 384
x = torch.randn(B, Cin, H, W, device=dev)
w = torch.randn(Cout, Cin, 1, 1, device=dev) / Cin ** 0.5
y = torch.empty(B, Cout, H, W, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
for _ in range(3):
    rc = ctx.L.irsde_conv2d_ex(ctx.h, 1, p(x), p(w), None, None, p(y), B, Cin, H, W, Cout, 1, 1, 1, 0, 0, 0, int(sys.argv[1]) if len(sys.argv) > 1 else 1,
                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    L.check(rc, ctx.h)
torch.cuda.synchronize()
print("ok")
