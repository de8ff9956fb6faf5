"""MN-major (channel-contiguous) tcgen05 operand probe: D[co,ci] = sum_p A[p][co] B[p][ci]."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pidnet_b200 import _lib
lib = _lib.load_probe()
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
a = torch.randn(64, 128, generator=g).to(torch.bfloat16).to(dev)
b = torch.randn(64, 64, generator=g).to(torch.bfloat16).to(dev)
ref = a.float().t() @ b.float()
out = torch.empty(128, 64, device=dev)
for lbo, sbo in [(8192, 1024), (1024, 8192), (8192, 2048), (16, 1024), (128, 1024)]:
    out.fill_(float('nan'))
    _lib.check(lib.pidnet_probe_mn(None, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), lbo, sbo, C.c_void_p(out.data_ptr())))
    err = (out - ref).abs()
    print(f'lbo={lbo} sbo={sbo}: max err {err.max().item():.4g}; rows 0-63 err {err[:64].max().item():.4g}, rows 64-127 err {err[64:].max().item():.4g} (scale {ref.abs().max().item():.3g})')
