"""CTA-pair (tcgen05 cta_group::2) probes: operand-split convention + cycles per M256 x N x K16 MMA."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pidnet_b200 import _lib
lib = _lib.load_probe()
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
a = torch.randn(256, 64, generator=g).bfloat16().to(dev)
b = torch.randn(64, 64, generator=g).bfloat16().to(dev)
ref = a.float() @ b.float().t()
for swap in (0, 1):
    out = torch.zeros(256, 64, device=dev)
    _lib.check(lib.pidnet_probe_pair(None, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), swap, C.c_void_p(out.data_ptr())))
    err = (out - ref).abs().max().item()
    ref_sw = torch.cat([ref[:, 32:], ref[:, :32]], 1)
    err_sw = (out - ref_sw).abs().max().item()
    print(f'swap_b={swap}: max|D - A B^T| = {err:.3g}; vs column-halves-swapped reference = {err_sw:.3g}')
for pairs in (1, 74):
    for N in (32, 64, 128):   # (N = 256 would need 9 x 16 KB of B per CTA on top of A: over the smem limit of this probe)
        for distinct in (0, 1):
            out = torch.zeros(pairs, dtype=torch.int64, device=dev)
            iters = 2048
            for _ in range(2):
                _lib.check(lib.pidnet_probe_mma_rate_pair(None, N, iters, distinct, pairs, C.c_void_p(out.data_ptr())))
            cyc = out.float().mean().item() / (iters * 4)
            print(f'pairs={pairs:3d} M=256 N={N:3d} distinct_operands={distinct}: {cyc:6.1f} cycles/MMA (tensor floor {128 * N / 256:.0f}, '
                  f'per-SM smem operand wavefronts {(128 + N // 2) * 32 // 128})')
