#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[4]): PIDNet-S, fwd + OhemCrossEntropy/BondaryLoss + bwd on
synthetic 12 x 3 x 1024 x 1024 crops per GPU, NCCL all-reduce of the flat fp32 gradient when WORLD_SIZE > 1.
Prints one JSON line (rank 0).  `torchrun --nproc-per-node N tools/bench_train.py` for N > 1."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# (and at every higher level), so route NCCL's log to stderr and drop a bare VERSION request
if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
    del os.environ['NCCL_DEBUG']
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
import torch
import torch.distributed as dist
from pidnet_b200 import PIDNet, OhemCrossEntropy, BondaryLoss
from pidnet_b200.criterion import FusedCriterion
from pidnet_b200.train import EngineTrainer

CW = [0.8373, 0.918, 0.866, 1.0345, 1.0166, 0.9969, 0.9754, 1.0489, 0.8786, 1.0023, 0.9539, 0.9843, 1.1116, 0.9037, 1.0865, 1.0955, 1.0865, 1.1529, 1.0507]

def main():
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    steps, warmup = int(os.environ.get('STEPS', 10)), int(os.environ.get('WARMUP', 3))
    B, H, W = int(os.environ.get('BATCH', 12)), int(os.environ.get('HEIGHT', 1024)), int(os.environ.get('WIDTH', 1024))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    model = PIDNet(m=2, n=3, num_classes=19, planes=32, ppm_planes=96, head_planes=128, augment=True).to(dev).train()
    tr = EngineTrainer(model)
    if os.environ.get('GRAPH', '1') == '0':
        tr.set_option('use_graph', 0)
    if os.environ.get('OVERLAP', '1') == '0':
        tr.set_option('overlap_wgrad', 0)
    weight = torch.tensor(CW)
    crit = FusedCriterion(OhemCrossEntropy(255, 0.9, 131072, weight), BondaryLoss())
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(B, 3, H, W, generator=g).to(dev)
    labels = torch.randint(0, 19, (B, H, W), generator=g)
    labels[:, :32, :] = 255
    labels = labels.to(dev)
    bd = (torch.rand(B, H, W, generator=g) > 0.9).float().to(dev)
    def one():
        out12, _ = tr.step(x, labels, bd, weight, crit.cfg, backward=True, want_logits=False)
        tr.allreduce_gradients()
        return out12
    for _ in range(warmup): one()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): out12 = one()
    e1.record()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if os.environ.get('QUICK'):      # profiler runs: one line, no extra phases
        if rank == 0:
            print(json.dumps(dict(ms_per_step=ms, loss=float(out12[0]))), flush=True)
        return
    # phase breakdown on rank 0 (forward+criterion only vs full)
    ef0, ef1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ef0.record()
    for _ in range(3): tr.step(x, labels, bd, weight, crit.cfg, backward=False, want_logits=False)
    ef1.record(); torch.cuda.synchronize()
    fwd_ms = ef0.elapsed_time(ef1) / 3
    # full iteration incl. the fused optimizer step (row f3): step + all-reduce + pidnet_sgd_step
    from pidnet_b200 import FusedSGD, adjust_learning_rate
    opt = FusedSGD(tr, lr=0.01, momentum=0.9, weight_decay=5e-4)
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record()
    for i in range(steps):
        one(); opt.step(); adjust_learning_rate(opt, 0.01, 1000, i)
    es1.record(); torch.cuda.synchronize()
    sgd_ms = es0.elapsed_time(es1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    import ctypes as C
    f, b = C.c_int(), C.c_int()
    tr.lib.pidnet_train_num_launches(tr.h, C.byref(f), C.byref(b))
    flops = 3 * 50.295e9 * B    # fwd (augment) 50.295 GFLOP / image at 1024x1024, fwd+bwd = 3x
    line = dict(metric='pidnet_s_train_1024x1024_images_per_sec', value=world * B * 1e3 / ms, unit='img/s', n_gpus=world,
                steps=steps, warmup=warmup, ms_per_step=ms, fwd_plus_criterion_ms=fwd_ms, bwd_ms=ms - fwd_ms,
                ms_per_step_with_sgd=sgd_ms,
                higher_is_better=True, scaling='weak', dtype='bf16 activations/gradients, fp32 master weights and weight gradients',
                data='synthetic', config=dict(workload=f'PIDNet-S train fwd + OHEM/boundary loss + bwd, {B}x3x{H}x{W} per GPU, NCCL all-reduce of {tr.n_param} fp32 gradients',
                batch_per_gpu=B), launches=dict(forward=f.value, backward=b.value), loss=float(out12[0]),
                conv_tflops=flops / (ms * 1e-3) / 1e12, grad_allreduce_bytes=4 * tr.n_param)
    if rank == 0 and os.environ.get('PROFILE'):
        buf = C.create_string_buffer(1 << 20)
        cm = C.c_float()
        p = lambda t: C.c_void_p(t.data_ptr())
        cw = weight.to(dev)
        tr.lib.pidnet_train_profile(tr.h, None, p(x), p(labels), p(bd), p(cw), C.byref(crit.cfg), buf, 1 << 20, C.byref(cm))
        rows = [l.split(' ', 3) for l in buf.value.decode().splitlines()]
        agg = {}
        for ph, ms_, kern, nm in rows:
            key = (ph, kern if kern != '-' else nm.split('.')[-1])
            a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += float(ms_)
        line['profile'] = dict(criterion_ms=cm.value, groups=sorted([dict(phase=k[0], kernel=k[1], launches=v[0], ms=round(v[1], 3)) for k, v in agg.items()], key=lambda d: -d['ms'])[:24])
        with open(os.environ['PROFILE'], 'w') as f:
            f.write(buf.value.decode())
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)

if __name__ == '__main__':
    main()
