"""Training-side boundary behaviour (reference: utils/utils.py:37-57 FullModel, utils/function.py:43-49 the loop,
tools/train.py:136-148): autograd hand-off in both forms, gradient-slot aliasing, stale-weight detection between
train and eval, label validation, and the C-ABI backward split used for the bucketed all-reduce."""
import ctypes as C

import pytest
import torch

from oracle import criterion_oracle as CO
from oracle import pidnet_oracle as O
from pidnet_b200 import BondaryLoss, FullModel, FusedSGD, OhemCrossEntropy, PIDNet, _lib

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device('cuda:0')


def _model(seed, dev, name='pidnet_s', ncls=19, train=True):
    cfg = O.config_for(name, ncls, True)
    model = PIDNet(m=cfg['m'], n=cfg['n'], num_classes=ncls, planes=cfg['planes'], ppm_planes=cfg['ppm_planes'],
                   head_planes=cfg['head_planes'], augment=True)
    model.load_state_dict(O.make_state_dict(cfg, seed, randomize_bn=False))
    model = model.to(dev)
    return model.train() if train else model.eval()


def _batch(N, H, W, seed, dev, ncls=19):
    x = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(seed)).to(dev)
    _, labels, bd = CO.synthetic_batch(N, ncls, H, W, seed + 1)
    return x, labels.to(dev), bd.to(dev)


def _full(model, dev, keep=3000):
    weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS)
    return FullModel(model, OhemCrossEntropy(255, 0.9, keep, weight), BondaryLoss(), return_outputs=False).to(dev).train()


def _flat(model):
    return torch.cat([p.grad.flatten() for p in model.parameters()]).clone()


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_eval_after_fused_sgd_sees_the_current_weights():
    """train (FusedSGD) -> eval -> train -> eval: the second validation must use the updated weights and running statistics
    (kernels rewrite the flat buffers without touching tensor._version)."""
    dev = _dev()
    model = _model(3, dev)
    full = _full(model, dev)
    opt = FusedSGD(full, lr=0.05, momentum=0.9, weight_decay=5e-4)
    x, labels, bd = _batch(4, 128, 256, 10, dev)
    xe = torch.randn(2, 3, 128, 256, generator=torch.Generator().manual_seed(77)).to(dev)
    seen, snaps = [], []
    for epoch in range(2):
        full.train()
        for _ in range(3):
            loss = full(x, labels, bd)[0].mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
        full.eval()
        snaps.append({k: v.detach().clone() for k, v in model.state_dict().items()})
        with torch.no_grad():
            got = model(xe)
            ref = O.pidnet_forward(snaps[-1], xe)
        torch.cuda.synchronize()
        err = O.rel_l2(got[1].cpu(), ref[1].cpu())
        print(f'[train/eval epoch {epoch}] eval logits vs oracle on the CURRENT state_dict: rel-L2 {err:.4g}')
        assert err < 8e-2, err
        seen.append(got[1].clone())
    with torch.no_grad():
        stale = O.pidnet_forward(snaps[0], xe)[1]
    err_stale = O.rel_l2(seen[1].cpu(), stale.cpu())
    print(f'[train/eval] second validation vs oracle on the FIRST epoch\'s weights: rel-L2 {err_stale:.4g}')
    assert err < 0.5 * err_stale, 'second validation ran on the first validation\'s weights (stale plan)'


def test_zero_grad_in_place_between_forward_and_backward():
    """Reference order: forward, model.zero_grad(), loss.backward() (utils/function.py:43-48).  With set_to_none=False the
    gradient slots are zeroed IN PLACE after the forward; the gradients must still arrive (they are computed in backward)."""
    dev = _dev()
    model = _model(4, dev)
    full = _full(model, dev)
    x, labels, bd = _batch(4, 128, 128, 20, dev)
    loss = full(x, labels, bd)[0].mean()
    full.zero_grad()
    loss.backward()
    g0 = _flat(model)
    assert float(g0.norm()) > 0
    # the .grad slots now alias the engine's flat buffer: zero them in place, as torch < 2.0 / set_to_none=False does
    loss = full(x, labels, bd)[0].mean()
    full.zero_grad(set_to_none=False)
    loss.backward()
    g1 = _flat(model)
    assert float(g1.norm()) > 0.5 * float(g0.norm()), 'zero_grad(set_to_none=False) wiped the gradients'
    # gradient accumulation over two backward passes (no zero_grad in between): the slots must hold the sum
    loss = full(x, labels, bd)[0].mean()
    loss.backward()
    g2 = _flat(model)
    loss = full(x, labels, bd)[0].mean()
    full.zero_grad()
    loss.backward()
    g3 = _flat(model)
    torch.cuda.synchronize()
    assert _rel(g2, g1 + g3) < 0.15, _rel(g2, g1 + g3)    # (same weights and batch: bf16 run-to-run noise only)
    assert float(g2.norm()) > 1.5 * float(g3.norm())


def test_loss_map_surface_in_train_mode():
    """FullModel(loss_map=True): the reference's return shapes ([1,N,H,W] loss, [N,H,W] loss_s); `.mean().backward()` delivers the
    same gradients as the 1-element form."""
    dev = _dev()
    x, labels, bd = _batch(2, 128, 128, 90, dev)
    weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS)
    grads, losses = [], []
    for lm in (False, True):
        model = _model(11, dev)
        full = FullModel(model, OhemCrossEntropy(255, 0.9, 3000, weight), BondaryLoss(), return_outputs=False, loss_map=lm).to(dev).train()
        loss, _, acc, ll = full(x, labels, bd)
        if lm:
            assert tuple(loss.shape) == (1, 2, 128, 128) and tuple(ll[0].shape) == (2, 128, 128)
        loss.mean().backward()
        grads.append(_flat(model)); losses.append(float(loss.mean()))
    torch.cuda.synchronize()
    assert abs(losses[0] - losses[1]) < 1e-4 * abs(losses[0])
    assert _rel(grads[1], grads[0]) < 0.05


def test_stale_or_repeated_backward_raises():
    dev = _dev()
    model = _model(5, dev)
    full = _full(model, dev)
    x, labels, bd = _batch(2, 128, 128, 30, dev)
    l1 = full(x, labels, bd)[0].mean()
    l2 = full(x, labels, bd)[0].mean()
    with pytest.raises(RuntimeError, match='earlier train-mode forward'):
        l1.backward()
    l2.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match='already back-propagated'):
        l2.backward()


def test_differentiable_train_forward_matches_fullmodel_gradients():
    """`outputs = model(inputs)` under autograd (utils/utils.py:39) + the reference's loss composition written in torch ops
    (the oracle restatement of FullModel.forward) must give the same parameter gradients as the fused FullModel path."""
    dev = _dev()
    x, labels, bd = _batch(4, 128, 256, 40, dev)
    weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS).to(dev)
    # (a) fused path
    ma = _model(6, dev)
    fa = _full(ma, dev, keep=3000)
    la = fa(x, labels, bd)[0].mean()
    la.backward()
    ga = _flat(ma)
    # (b) the engine as a plain differentiable module under a torch-side loss
    mb = _model(6, dev)
    outs = mb(x)
    assert isinstance(outs, list) and len(outs) == 3 and all(o.requires_grad for o in outs)
    losses, _, acc, _ = CO.full_model_forward(outs, labels, bd, weight, dict(ohem_keep=3000))
    lb = losses.mean()
    lb.backward()
    gb = _flat(mb)
    torch.cuda.synchronize()
    assert abs(float(la) - float(lb)) < 2e-3 * abs(float(lb)), (float(la), float(lb))
    r = _rel(ga, gb)
    cos = float(ga @ gb / (ga.norm() * gb.norm()))
    print(f'[differentiable forward] fused FullModel vs model(x) + torch loss: gradient rel-L2 {r:.3g}, cosine {cos:.5f}')
    assert cos > 0.99 and r < 0.15, (r, cos)     # two engine runs: bf16 run-to-run noise (atomics order) only
    # torch.optim.SGD on the published gradients moves the weights
    opt = torch.optim.SGD(mb.parameters(), lr=0.01, momentum=0.9)
    w0 = mb.final_layer.conv2.weight.detach().clone()
    opt.step()
    assert not torch.equal(w0, mb.final_layer.conv2.weight.detach())
    # partial use of the outputs: unused heads receive zero gradient
    outs = mb(x)
    outs[1].square().mean().backward()
    assert all(torch.isfinite(p.grad).all() for p in mb.parameters())


def test_out_of_range_labels_are_reported_not_dereferenced():
    """OhemCrossEntropy defaults to ignore_label=-1 (as configs/default.py); Cityscapes labels carry 255.  The reference's
    gather faults on the device; here the pixels are dropped, the loss is NaN and the trainer raises at the next step."""
    dev = _dev()
    model = _model(7, dev)
    weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS)
    full = FullModel(model, OhemCrossEntropy(-1, 0.9, 2000, weight), BondaryLoss(), return_outputs=False).to(dev).train()
    x, labels, bd = _batch(2, 128, 128, 50, dev)
    assert int((labels == 255).sum()) > 0
    loss = full(x, labels, bd)[0]
    torch.cuda.synchronize()
    assert loss.isnan().all()
    with pytest.raises(RuntimeError, match='neither ignore_label'):
        full.check_valid()
    # eval-mode criterion call reports through the output vector
    full.eval()
    with torch.no_grad():
        out, _ = full._crit(model(x), labels, bd)
    assert float(out[12]) == float((labels == 255).sum())


def test_empty_ohem_set_raises_like_the_reference():
    dev = _dev()
    model = _model(8, dev)
    full = _full(model, dev)
    x, labels, bd = _batch(2, 128, 128, 60, dev)
    labels = torch.full_like(labels, 255)
    loss = full(x, labels, bd)[0].mean()
    loss.backward()
    torch.cuda.synchronize()
    assert loss.isnan()
    g = _flat(model)
    assert torch.isfinite(g).all(), 'an empty OHEM selection leaked NaN into the parameter gradients'
    with pytest.raises(IndexError):
        full.check_valid()


def test_detached_storage_is_detected():
    dev = _dev()
    model = _model(9, dev)
    full = _full(model, dev)
    x, labels, bd = _batch(2, 128, 128, 70, dev)
    full(x, labels, bd)[0].mean().backward()
    model.double().float()        # re-seats every p.data
    with pytest.raises(RuntimeError, match='storage of'):
        full(x, labels, bd)


def test_segmented_backward_equals_whole_backward_and_ranges_partition_the_gradient():
    """pidnet_train_backward in ranges (the hook for the bucketed all-reduce) == the single-call backward; the ranges reported
    final after each segment are disjoint and cover every parameter gradient."""
    dev = _dev()
    x, labels, bd = _batch(4, 128, 256, 80, dev)
    cc = _lib.CriterionCfg(ignore_label=255, ohem_thres=0.9, ohem_keep=3000, bd_threshold=0.8,
                           balance_weight_aux=0.4, balance_weight_main=1.0, sb_weight=1.0, coeff_bce=20.0)
    grads = {}
    for mode in ('whole', 'segments', 'whole2', 'segments_graph'):
        model = _model(10, dev)
        tr = model.engine_trainer()
        tr.set_option('use_graph', 1 if mode == 'segments_graph' else 0)
        for it in range(3):        # later iterations exercise the steady state (CUDA-graph replay in the last mode)
            tr.step(x, labels, bd, None, cc, backward=2, want_logits=False)
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            if not mode.startswith('segments'):
                _lib.check(tr.lib.pidnet_train_backward(tr.h, stream, C.c_void_p(x.data_ptr()), None, None, None, -1))
            else:
                for s in range(tr.lib.pidnet_train_num_segments(tr.h)):
                    _lib.check(tr.lib.pidnet_train_backward(tr.h, stream, C.c_void_p(x.data_ptr()), None, None, None, s))
        torch.cuda.synchronize()
        grads[mode] = tr.flat_grad.clone()
        if mode == 'segments':
            ranges = tr.segment_ranges()
            assert len(ranges) == tr.lib.pidnet_train_num_segments(tr.h) >= 2
            cover = torch.zeros(tr.flat_grad.numel(), dtype=torch.int32)
            for seg in ranges:
                for b, e in seg:
                    assert 0 <= b < e <= tr.n_param
                    cover[b:e] += 1
            assert int(cover.max()) == 1, 'a gradient range is reported final twice'
            off = 0
            for p in model.parameters():
                assert int(cover[off:off + p.numel()].min()) == 1, 'a parameter gradient is never reported final'
                off += (p.numel() + 3) // 4 * 4
            # late ranges hold the small high-resolution layers: most bytes are final before the last segment starts
            last_bytes = sum(e - b for b, e in ranges[-1])
            assert last_bytes < 0.2 * tr.n_param, (last_bytes, tr.n_param)
    noise = _rel(grads['whole2'], grads['whole'])
    diff = _rel(grads['segments'], grads['whole'])
    diffg = _rel(grads['segments_graph'], grads['whole'])
    print(f'[segmented backward] segments vs whole {diff:.3e} (graph replay {diffg:.3e}); whole vs whole (run-to-run) {noise:.3e}')
    # (gradients are not bit-reproducible run to run: the BatchNorm sums are fp64 atomics of fp32 partials, and a flipped bf16
    # rounding is amplified by the batch-norm backward chain -- same bound as tests/test_train_gpu.py's graph-vs-eager check)
    assert diff <= max(3 * noise, 0.05) and diffg <= max(3 * noise, 0.05), (diff, diffg, noise)


def test_two_rank_nccl_gradient_is_the_mean_of_the_shard_gradients():
    """torchrun x2 over NCCL (tests/nccl_grad_worker.py): the bucketed, overlapped all-reduce must deliver exactly the mean of the
    two ranks' local engine gradients, and that mean must agree with the mean of the two fp32-oracle shard gradients
    (reference: nn.DataParallel reduce-add + losses.mean() over replicas, tools/train.py:136 / utils/function.py:44)."""
    import os, subprocess, sys
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run under `gpurun --gpus 2`)')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(root, 'tests', 'nccl_grad_worker.py')]
    r = subprocess.run(cmd, cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and 'NCCL_GRAD_OK' in r.stdout
