"""Microbenchmark of the HBM-bound fused kernels at the PIDNet-S 32x1024x2048 shapes (through the C ABI op exports)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pidnet_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
N = int(os.environ.get('N', 32))
def t(*shape): return (torch.randn(*shape, device=dev) * 0.5).bfloat16()
def P(x): return C.c_void_p(x.data_ptr()) if x is not None else None
def timeit(name, fn, nbytes, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f'{name:34s} {ms:.4f} ms  {nbytes / ms / 1e6:7.0f} GB/s ({nbytes / 1e6:.0f} MB)')
only = sys.argv[1] if len(sys.argv) > 1 else ''
cases = []
for name, (H, W, Cc, h, w) in {'pag3': (128, 256, 64, 64, 128), 'pag4': (128, 256, 64, 32, 64)}.items():
    x, low, out = t(N, H, W, Cc), t(N, h, w, 2 * Cc + 8), t(N, H, W, Cc)
    cases.append((name + '.fuse', lambda x=x, low=low, out=out, H=H, W=W, Cc=Cc, h=h, w=w: _lib.check(
        lib.pidnet_op_pag(None, P(x), P(low), P(out), N, H, W, Cc, h, w, 1)), 2 * (x.numel() + out.numel() + low.numel())))
for name, (H, W, Cc, h, w) in {'diff3.upadd': (128, 256, 64, 64, 128), 'diff4.upadd': (128, 256, 128, 32, 64)}.items():
    a, b, out = t(N, H, W, Cc), t(N, h, w, Cc), t(N, H, W, Cc)
    cases.append((name, lambda a=a, b=b, out=out, H=H, W=W, Cc=Cc, h=h, w=w: _lib.check(
        lib.pidnet_op_upadd(None, P(a), P(b), P(out), N, H, W, Cc, h, w, None, None, 1)), 2 * (a.numel() + b.numel() + out.numel())))
p, il, d, out = t(N, 128, 256, 128), t(N, 16, 32, 128), t(N, 128, 256, 128), t(N, 128, 256, 256)
cases.append(('dfm.uv', lambda: _lib.check(lib.pidnet_op_lightbag(None, P(p), P(il), P(d), P(out), N, 128, 256, 128, 16, 32)),
              2 * (p.numel() + il.numel() + d.numel() + out.numel())))
for name, fn, nb in cases:
    if only in name:
        timeit(name, fn, nb)
