"""How far apart do 50-step loss curves of the SAME training run land when only rounding differs?  Trajectories: fp32 oracle;
fp32 oracle from initial weights perturbed by one bf16 ulp-ish relative noise; bf16-storage-emulated oracle; the engine (twice).
Used to set the tolerance of tests/test_baseline_parity_gpu.py::test_fifty_step_loss_curve_tracks_the_fp32_reference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import criterion_oracle as CO, pidnet_oracle as O
from pidnet_b200 import BondaryLoss, FullModel, FusedSGD, OhemCrossEntropy, PIDNet
from tools import train_check
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_baseline_parity_gpu import synth_batch, NCLS


def oracle_run(sd0, batches, dev, weight, keep, lr, perturb=0.0, emulate=False):
    saved = train_check.emulate_bf16_storage() if emulate else None
    try:
        sd = {k: v.clone().to(dev) for k, v in sd0.items()}
        if perturb:
            g = torch.Generator(device='cpu').manual_seed(1)
            for k, v in sd.items():
                if v.dtype.is_floating_point and 'running_' not in k:
                    v.mul_(1 + perturb * torch.randn(v.shape, generator=g).to(dev))
        params = [v for k, v in sd.items() if v.dtype.is_floating_point and 'running_' not in k]
        for p in params:
            p.requires_grad_(True)
        opt = torch.optim.SGD(params, lr=lr, momentum=0.9, weight_decay=5e-4)
        out = []
        for x, y, bd in batches:
            outs = O.pidnet_forward(sd, x.to(dev), training=True)
            losses, _, _, _ = CO.full_model_forward(list(outs), y.to(dev), bd.to(dev), weight.to(dev), dict(ohem_keep=keep))
            loss = losses.mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            out.append(float(loss))
        return out
    finally:
        if saved:
            O._conv, O._bn = saved


def engine_run(cfg, sd0, batches, dev, weight, keep, lr):
    model = PIDNet(m=cfg['m'], n=cfg['n'], num_classes=NCLS, planes=cfg['planes'], ppm_planes=cfg['ppm_planes'],
                   head_planes=cfg['head_planes'], augment=True)
    model.load_state_dict(sd0)
    full = FullModel(model, OhemCrossEntropy(255, 0.9, keep, weight), BondaryLoss(), return_outputs=False).to(dev).train()
    opt = FusedSGD(full, lr=lr, momentum=0.9, weight_decay=5e-4)
    out = []
    for x, y, bd in batches:
        loss = full(x.to(dev), y.to(dev), bd.to(dev))[0].mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        out.append(float(loss))
    return out


def main():
    dev = torch.device('cuda:0')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS)
    for (N, H, W, steps, lr, keep) in [(4, 256, 512, 50, 0.01, 20000), (8, 256, 512, 50, 0.01, 40000), (8, 256, 512, 100, 0.003, 40000)]:
        cfg = O.config_for('pidnet_s', NCLS, True)
        sd0 = O.make_state_dict(cfg, 21, randomize_bn=False)
        gen = torch.Generator().manual_seed(5)
        palette = torch.randn(NCLS, 3, generator=gen)
        batches = [synth_batch(N, H, W, gen, palette, 'cpu') for _ in range(steps)]
        runs = {
            'fp32': oracle_run(sd0, batches, dev, weight, keep, lr),
            'fp32 perturbed 2^-9': oracle_run(sd0, batches, dev, weight, keep, lr, perturb=2.0 ** -9),
            'fp32 perturbed 2^-12': oracle_run(sd0, batches, dev, weight, keep, lr, perturb=2.0 ** -12),
            'bf16-emulated oracle': oracle_run(sd0, batches, dev, weight, keep, lr, emulate=True),
            'engine': engine_run(cfg, sd0, batches, dev, weight, keep, lr),
            'engine again': engine_run(cfg, sd0, batches, dev, weight, keep, lr),
        }
        print(f'--- N={N} {H}x{W} steps={steps} lr={lr} keep={keep}')
        for k, v in runs.items():
            tail = sum(v[-10:]) / 10
            print(f'{k:24s} first {v[0]:.3f} tail10 {tail:.4f}  ({100 * (tail / (sum(runs["fp32"][-10:]) / 10) - 1):+.1f} % vs fp32)  '
                  + ' '.join(f'{x:.2f}' for x in v[::max(1, steps // 10)]))


if __name__ == '__main__':
    main()
