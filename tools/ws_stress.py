"""Persistent-kernel stress: many tiles per CTA for the weight-stationary kernels (both modes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from tests import util as U
dev = torch.device('cuda:0')
cases = [  # N,H,W,Cin,Cout,k,res
    (8, 128, 256, 64, 64, 3, True), (8, 128, 256, 64, 64, 3, False),
    (8, 64, 128, 128, 128, 3, False), (8, 64, 128, 128, 128, 3, True),
    (4, 256, 512, 32, 32, 3, True), (8, 128, 256, 128, 128, 1, True), (8, 128, 256, 256, 128, 1, False),
]
for (N, H, W, Ci, Co, k, res) in cases:
    g = torch.Generator().manual_seed(1)
    x = U.bf16r(torch.randn(N, Ci, H, W, generator=g)).to(dev)
    w = U.bf16r(torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5)
    b = torch.randn(Co, generator=g)
    r = U.bf16r(torch.randn(N, Co, H, W, generator=g)).to(dev) if res else None
    try:
        out = U.op_conv2d(U.to_nhwc_bf16(x), w, b, 1, 1, U.to_nhwc_bf16(r) if res else None, True, False, 0)
        ref = F.conv2d(x, w.to(dev), b.to(dev), 1, k // 2)
        if res: ref = ref + r
        ref = ref.relu()
        err = (U.from_nhwc(out) - ref).abs().max().item() / ref.abs().max().item()
        print((N, H, W, Ci, Co, k, res), 'ok rel max err %.4g' % err, flush=True)
    except Exception as e:
        print((N, H, W, Ci, Co, k, res), 'FAILED', str(e)[:200], flush=True)
        break
