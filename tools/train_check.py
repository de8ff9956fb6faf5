"""Training-step parity, stage-local: every oracle stage is re-run with torch autograd (fp32, train-mode BN) from
the ENGINE's own stage inputs and the ENGINE's own upstream gradient; the stage's parameter gradients and its
contributions to the input gradients must match what the engine produced.  (End to end the comparison is
meaningless on random-init weights: train-mode BatchNorm makes the bf16 forward chaotic -- a bf16-emulated
oracle deviates from the fp32 oracle exactly as the engine does, 0.5 % at conv1 growing to 10 % at layer5.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pidnet_oracle as O, criterion_oracle as CO
from pidnet_b200 import PIDNet, OhemCrossEntropy, BondaryLoss
from pidnet_b200.criterion import FusedCriterion
from pidnet_b200.train import EngineTrainer


def _bf(t):
    return t.to(torch.bfloat16).float()


def _ste(t):
    # value rounded to bf16 in the forward pass, gradient rounded to bf16 in the backward pass
    y = t + (_bf(t) - t).detach()
    if y.requires_grad:
        y.register_hook(_bf)
    return y


def emulate_bf16_storage():
    """Make the oracle round every stored tensor (conv outputs, BN outputs, their gradients) to bf16, like the engine."""
    import torch.nn.functional as F
    oc, ob = O._conv, O._bn

    def conv_e(c, x, name, stride=1, padding=0, groups=1):
        w = c.sd[name + '.weight']
        w = w + (_bf(w) - w).detach()      # bf16 operand copy of the fp32 master weight; its gradient stays fp32 (engine: wgrad)
        return _ste(F.conv2d(x, w, c.sd.get(name + '.bias'), stride, padding, 1, groups))

    def bn_e(c, x, name):
        return _ste(ob(c, x, name))
    O._conv, O._bn = conv_e, bn_e
    return oc, ob


def run(name='pidnet_s', ncls=19, N=4, H=256, W=256, keep=4000, seed=5, verbose=True, emulate=False, also_emulated=False):
    """also_emulated: additionally run the stage-local oracle with bf16 STORAGE emulation (every conv / BN output and its gradient
    rounded to bf16, bf16 weight operands -- what any bf16 implementation of the reference does) and append, per parameter,
    the emulated oracle's own distance from the fp32 oracle: the yardstick the engine's distance is asserted against."""
    dev = torch.device('cuda:0')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = O.config_for(name, ncls, True)
    sd = O.make_state_dict(cfg, seed)
    model = PIDNet(m=cfg['m'], n=cfg['n'], num_classes=ncls, planes=cfg['planes'], ppm_planes=cfg['ppm_planes'],
                   head_planes=cfg['head_planes'], augment=True)
    model.load_state_dict(sd)
    model = model.to(dev).train()
    x = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(3)).to(dev)
    _, labels, bd = CO.synthetic_batch(N, ncls, H, W, 9)
    labels, bd = labels.to(dev), bd.to(dev)
    weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS[:ncls]).to(dev)
    tr = EngineTrainer(model)
    crit = FusedCriterion(OhemCrossEntropy(255, 0.9, keep, weight), BondaryLoss())
    out12, eouts = tr.step(x, labels, bd, weight, crit.cfg, backward=True)
    torch.cuda.synchronize()
    # ---- criterion on the engine's logits (fp32 oracle) -> reference dlogits
    lg = [t.detach().clone().requires_grad_(True) for t in eouts]
    losses, _, acc, ll = CO.full_model_forward(lg, labels, bd, weight, dict(ohem_keep=keep))
    losses.mean().backward()
    res = dict(loss=(float(out12[0]), float(losses.mean())), acc=(float(out12[3]), float(acc)))
    up = {'out_p': lg[0].grad, 'out': lg[1].grad, 'out_d': lg[2].grad}
    eng_out = {'out_p': eouts[0], 'out': eouts[1], 'out_d': eouts[2]}
    # ---- stage-local autograd
    stages = O._stages(sd)
    names = [s[0] for s in stages]
    ev = {'x': x}
    eg = {}
    for nm in names:
        if nm in eng_out:
            ev[nm] = eng_out[nm].detach()
        else:
            ev[nm] = tr.debug_tensor(nm).to(dev)
            eg[nm] = tr.debug_tensor(nm, grad=True).to(dev)
    emu_pg = _stage_param_grads(sd, ev, eg, up, H, W, dev) if also_emulated else None
    saved = emulate_bf16_storage() if emulate else None
    ref_g = {nm: torch.zeros_like(ev[nm]) for nm in names}
    ref_pg = {}
    ref_run = {}
    fwd_err = {}
    for nm, ins, fn in stages:
        psd = {k: v.detach().clone().to(dev) for k, v in sd.items()}
        for k, v in psd.items():
            if v.dtype.is_floating_point and 'running_' not in k:
                v.requires_grad_(True)
        c = O._Ctx(psd, True)
        c.size8 = (H // 8, W // 8)
        inputs = [ev[i].detach().clone().requires_grad_(i != 'x') for i in ins]
        out = fn(c, *inputs)
        fwd_err[nm] = O.rel_l2(ev[nm].cpu(), out.detach().cpu())
        out.backward(up[nm] if nm in up else eg[nm])
        for i, t in zip(ins, inputs):
            if i != 'x' and t.grad is not None:
                ref_g[i] += t.grad
        for k, v in psd.items():
            if v.requires_grad and v.grad is not None and float(v.grad.abs().sum()) > 0:
                ref_pg[k] = ref_pg.get(k, 0) + v.grad
            if 'running_' in k and not torch.equal(v.cpu(), sd[k]):
                ref_run[k] = v.detach().clone()
    if saved:
        O._conv, O._bn = saved
    # running statistics: every BN belongs to exactly one stage; compare the oracle's update (from engine inputs)
    run_err = {}
    rows_t = []
    for nm in names:
        if nm in eg:
            a, b = eg[nm].double().flatten(), ref_g[nm].double().flatten()
            rows_t.append((nm, float((a - b).norm() / (b.norm() + 1e-30)), float(a @ b / (a.norm() * b.norm() + 1e-30))))
    msd = model.state_dict()
    for k, v in ref_run.items():
        run_err[k] = float((msd[k].to(dev) - v).abs().max() / (v.abs().max() + 1e-6))
    rows_p = []
    for k, gr in ref_pg.items():
        a, b = tr.grad_views[k].double().flatten(), gr.double().flatten()
        rows_p.append((k, float((a - b).norm() / (b.norm() + 1e-30)), float(a @ b / (a.norm() * b.norm() + 1e-30)), float(b.norm())))
    if emu_pg is not None:
        rows_p = [r + (float((emu_pg[r[0]].double().flatten() - ref_pg[r[0]].double().flatten()).norm() / (r[3] + 1e-30))
                       if r[0] in emu_pg else float('nan'),) for r in rows_p]
    if verbose:
        print('loss engine/ref', res['loss'], 'acc', res['acc'])
        print('stage-local forward rel-L2 :', ' '.join('%s=%.3g' % kv for kv in fwd_err.items()))
        print('stage-input gradients (rel-L2, cos):', ' '.join('%s=%.3g/%.4f' % r for r in rows_t))
        rows_p.sort(key=lambda r: -r[1])
        print('parameter gradients, worst 30 of %d (rel-L2, cos, |ref|):' % len(rows_p))
        for r in rows_p[:30]:
            print('   %-40s %.4g  %.5f  %.3g' % r[:4])
        import statistics
        print('median param-grad rel-L2 %.4g' % statistics.median([r[1] for r in rows_p]))
        print('running stats updated: %d, max rel err %.3g' % (len(run_err), max(run_err.values())))
    return res, fwd_err, rows_t, rows_p, run_err


def _stage_param_grads(sd, ev, eg, up, H, W, dev):
    """Stage-local parameter gradients of the bf16-storage-emulated oracle (same engine stage inputs / upstream gradients)."""
    saved = emulate_bf16_storage()
    try:
        pg = {}
        for nm, ins, fn in O._stages(sd):
            psd = {k: v.detach().clone().to(dev) for k, v in sd.items()}
            for k, v in psd.items():
                if v.dtype.is_floating_point and 'running_' not in k:
                    v.requires_grad_(True)
            c = O._Ctx(psd, True)
            c.size8 = (H // 8, W // 8)
            inputs = [ev[i].detach().clone().requires_grad_(i != 'x') for i in ins]
            out = fn(c, *inputs)
            out.backward(up[nm] if nm in up else eg[nm])
            for k, v in psd.items():
                if v.requires_grad and v.grad is not None and float(v.grad.abs().sum()) > 0:
                    pg[k] = pg.get(k, 0) + v.grad
    finally:
        O._conv, O._bn = saved
    return pg


if __name__ == '__main__':
    run(sys.argv[1] if len(sys.argv) > 1 else 'pidnet_s', emulate=os.environ.get('EMUL', '0') == '1')
