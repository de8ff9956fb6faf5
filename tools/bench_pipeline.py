#!/usr/bin/env python
"""Device-time breakdown of the camera-frame pipeline (PIDNet.segment): fp32 forward vs uint8 forward vs the fused
x8 upsample + argmax, PIDNet-S, batch 32, 1024x2048."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pidnet_b200 import get_pred_model
from pidnet_b200.postprocess import upsample_argmax

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    B, H, W = int(os.environ.get('BATCH', 32)), 1024, 2048
    dev = torch.device('cuda:0')
    model = get_pred_model('pidnet_s', 19).to(dev).eval()
    x = torch.randn(B, 3, H, W, device=dev)
    fr = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
    with torch.no_grad():
        logits = model(x)
        res = dict(batch=B, forward_fp32_ms=timeit(lambda: model(x)), forward_u8_ms=timeit(lambda: model.forward_u8(fr)),
                   upsample_argmax_ms=timeit(lambda: upsample_argmax(logits, (H, W))), segment_ms=timeit(lambda: model.segment(fr)))
    print(json.dumps(res))

if __name__ == '__main__':
    main()
