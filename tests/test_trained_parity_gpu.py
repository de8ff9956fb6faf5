"""North-star parity gate on TRAINED-LIKE weights: bf16 engine vs fp32 oracle logits within 2e-2
relative error and >= 99.9 % argmax agreement (BASELINE.json north_star).

No checkpoints or datasets exist offline, and bf16 cannot reach 99.9 % on random-init weights for ANY
implementation, so the weights are produced here: PIDNet-S is trained for a few
hundred SGD steps on a synthetic blocky-label task with the ORACLE's own forward in train mode (torch
autograd on the GPU -- test infrastructure only), at the evaluation geometry, then frozen and evaluated
on fresh in-distribution images through (a) the oracle in fp32 and (b) the engine.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import pidnet_oracle as O
from pidnet_b200 import PIDNet

pytestmark = pytest.mark.gpu

NCLS, H, W = 19, 256, 512
REL_TOL = 2e-2          # north star: bf16 within 2e-2 relative error
ARGMAX_TOL = 0.999      # north star: argmax agreement >= 99.9 % of pixels


def synth_batch(n, gen, palette, dev):
    """Blocky random label maps; image = per-class colour + noise."""
    coarse = torch.randint(0, NCLS, (n, 1, H // 32, W // 32), generator=gen).float()
    labels = F.interpolate(coarse, size=(H, W), mode='nearest').long()[:, 0]
    img = palette[labels].permute(0, 3, 1, 2).contiguous()
    img = img + 0.25 * torch.randn(img.shape, generator=gen)
    return img.to(dev), labels.to(dev)


def train_state_dict(dev, steps=400, batch=4):
    cfg = O.config_for('pidnet_s', NCLS, True)
    sd = {k: v.to(dev) for k, v in O.make_state_dict(cfg, 7, randomize_bn=False).items()}
    params = [v for k, v in sd.items() if v.dtype.is_floating_point and 'running_' not in k]
    for p in params:
        p.requires_grad_(True)
    opt = torch.optim.SGD(params, lr=0.01, momentum=0.9, weight_decay=5e-4)
    gen = torch.Generator().manual_seed(123)
    palette = torch.randn(NCLS, 3, generator=gen)
    for it in range(steps):
        x, y = synth_batch(batch, gen, palette, dev)
        y8 = y[:, 4::8, 4::8]
        aux_p, main, aux_d = O.pidnet_forward(sd, x, training=True)
        loss = F.cross_entropy(main, y8) + 0.4 * F.cross_entropy(aux_p, y8)
        # boundary head: label edges at 1/8 resolution
        edge = ((y8[:, 1:, :-1] != y8[:, :-1, :-1]) | (y8[:, :-1, 1:] != y8[:, :-1, :-1])).float()
        loss = loss + F.binary_cross_entropy_with_logits(aux_d[:, 0, :-1, :-1], edge)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    for p in params:
        p.requires_grad_(False)
    return cfg, {k: v.detach() for k, v in sd.items()}, palette, float(loss.detach())


def test_trained_weights_meet_north_star_parity():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    dev = torch.device('cuda:0')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    cfg, sd, palette, last_loss = train_state_dict(dev)
    gen = torch.Generator().manual_seed(999)
    x, y = synth_batch(4, gen, palette, dev)
    with torch.no_grad():
        ref = O.pidnet_forward(sd, x)                         # fp32 oracle (cuDNN, TF32 off)
        ref_cpu = O.pidnet_forward({k: v.cpu() for k, v in sd.items()}, x[:1].cpu())
    model = PIDNet(m=cfg['m'], n=cfg['n'], num_classes=NCLS, planes=cfg['planes'], ppm_planes=cfg['ppm_planes'],
                   head_planes=cfg['head_planes'], augment=True)
    model.load_state_dict({k: v.cpu() for k, v in sd.items()})
    model = model.to(dev).eval()
    with torch.no_grad():
        got = model(x)
    torch.cuda.synchronize()
    acc = float((ref[1].argmax(1) == y[:, 4::8, 4::8]).float().mean())
    print(f'\\n[trained parity] final train loss {last_loss:.4f}, oracle pixel accuracy on fresh images {acc:.4f}')
    # the GPU fp32 oracle must itself agree with the CPU fp32 oracle (different conv algorithms, same math)
    assert O.rel_l2(ref[1][:1].cpu(), ref_cpu[1]) < 1e-4
    names = ('x_extra_p', 'x_', 'x_extra_d')
    for nm, g, r in zip(names, got, ref):
        rel = O.rel_l2(g.cpu(), r.cpu())
        agree = O.argmax_agreement(g.cpu(), r.cpu()) if g.shape[1] > 1 else float(((g > 0) == (r > 0)).float().mean())
        print(f'[trained parity] {nm}: rel-L2 {rel:.4g}  argmax agreement {agree:.5f}')
        assert rel < REL_TOL, f'{nm}: rel-L2 {rel:.4g} >= {REL_TOL}'
        if nm == 'x_':
            assert agree >= ARGMAX_TOL, f'{nm}: argmax agreement {agree:.5f} < {ARGMAX_TOL}'
    assert acc > 0.9, 'synthetic training did not converge; parity weights are not trained-like'
