"""Run the halo-patch descriptor probe (pidnet_b200/csrc/probe.cu) for all nine taps."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pidnet_b200 import _lib
lib = _lib.load_probe()
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
x = torch.randn(18, 10, 64, generator=g).to(torch.bfloat16).to(dev)
w = torch.randn(64, 64, generator=g).to(torch.bfloat16).to(dev)
out = torch.empty(128, 64, device=dev)
for mode in (0, 1):
    for r in range(3):
        for s in range(3):
            out.fill_(float('nan'))
            _lib.check(lib.pidnet_probe_halo(None, C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), r, s, mode,
                                             C.c_void_p(out.data_ptr())))
            ref = (x[r:r + 16, s:s + 8].reshape(128, 64).float() @ w.float().t())
            err = (out - ref).abs().max().item()
            print(f'mode={mode} tap=({r},{s}) max abs err {err:.4g}  (ref scale {ref.abs().max().item():.3g})')
