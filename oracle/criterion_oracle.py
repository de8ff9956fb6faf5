"""CPU oracle for the training/validation criterion of the PIDNet path -- TEST INFRASTRUCTURE
(same rules as oracle/pidnet_oracle.py).

Restates, with the config values passed explicitly instead of read from the global yacs singleton:
  * FullModel.forward            /root/reference/utils/utils.py:37-57
  * OhemCrossEntropy             /root/reference/utils/criterion.py:43-99
  * weighted_bce / BondaryLoss   /root/reference/utils/criterion.py:102-132
Pinned against the live reference (imported with a yacs shim) in tests/test_criterion_oracle.py and
against golden vectors in tests/golden/criterion_*.npz (tools/make_golden_criterion.py).
"""
import torch
import torch.nn.functional as F

# values of the 8 shipped YAMLs; configs/default.py differs
DEFAULTS = dict(ignore_label=255, ohem_thres=0.9, ohem_keep=131072, balance_weights=(0.4, 1.0), sb_weights=1.0,
                align_corners=True, coeff_bce=20.0, bd_threshold=0.8)

CITYSCAPES_CLASS_WEIGHTS = [0.8373, 0.918, 0.866, 1.0345, 1.0166, 0.9969, 0.9754, 1.0489, 0.8786, 1.0023,
                            0.9539, 0.9843, 1.1116, 0.9037, 1.0865, 1.0955, 1.0865, 1.1529, 1.0507]   # datasets/cityscapes.py:55-59


def ce_none(score, target, weight, ignore_label):
    """nn.CrossEntropyLoss(weight, ignore_index, reduction='none')  (criterion.py:50-54)"""
    return F.cross_entropy(score, target, weight=weight, ignore_index=ignore_label, reduction='none')


def ohem_forward(score, target, weight, ignore_label, thres, min_kept):
    """criterion.py:63-78 (raises IndexError when no pixel is valid, like the reference)."""
    pred = F.softmax(score, dim=1)
    pixel_losses = ce_none(score, target, weight, ignore_label).contiguous().view(-1)
    mask = target.contiguous().view(-1) != ignore_label
    tmp_target = target.clone()
    tmp_target[tmp_target == ignore_label] = 0
    pred = pred.gather(1, tmp_target.unsqueeze(1))
    pred, ind = pred.contiguous().view(-1,)[mask].contiguous().sort()
    min_value = pred[min(max(1, min_kept), pred.numel() - 1)]
    threshold = max(min_value, thres)
    pixel_losses = pixel_losses[mask][ind]
    pixel_losses = pixel_losses[pred < threshold]
    return pixel_losses.mean()


def sem_loss(scores, target, weight, cfg):
    """OhemCrossEntropy.forward (criterion.py:80-99)."""
    if not isinstance(scores, (list, tuple)):
        scores = [scores]
    bw = cfg['balance_weights']
    args = (weight, cfg['ignore_label'])
    if len(bw) == len(scores):
        terms = [w * ce_none(x, target, *args) for w, x in zip(bw[:-1], scores[:-1])]
        terms.append(bw[-1] * ohem_forward(scores[-1], target, *args, cfg['ohem_thres'], cfg['ohem_keep']))
        return sum(terms)
    if len(scores) == 1:
        return cfg['sb_weights'] * ohem_forward(scores[0], target, *args, cfg['ohem_thres'], cfg['ohem_keep'])
    raise ValueError('lengths of prediction and target are not identical!')


def weighted_bce(bd_pre, target):
    """criterion.py:102-119"""
    log_p = bd_pre.permute(0, 2, 3, 1).contiguous().view(1, -1)
    target_t = target.view(1, -1)
    pos_index = (target_t == 1)
    neg_index = (target_t == 0)
    weight = torch.zeros_like(log_p)
    pos_num = pos_index.sum()
    neg_num = neg_index.sum()
    sum_num = pos_num + neg_num
    weight[pos_index] = neg_num * 1.0 / sum_num
    weight[neg_index] = pos_num * 1.0 / sum_num
    return F.binary_cross_entropy_with_logits(log_p, target_t, weight, reduction='mean')


def full_model_forward(outputs, labels, bd_gt, weight=None, cfg=None):
    """FullModel.forward AFTER `outputs = self.model(inputs)` (utils/utils.py:41-57).

    outputs: [x_extra_p, x_, x_extra_d] low-res logits.  Returns exactly what the reference returns:
    (loss.unsqueeze(0) [1,N,H,W], [up(x_extra_p), up(x_)], acc, [loss_s [N,H,W], loss_b scalar])."""
    cfg = dict(DEFAULTS, **(cfg or {}))
    outputs = list(outputs)
    h, w = labels.size(1), labels.size(2)
    if outputs[0].size(2) != h or outputs[0].size(3) != w:
        outputs = [F.interpolate(o, size=(h, w), mode='bilinear', align_corners=cfg['align_corners']) for o in outputs]
    _, preds = torch.max(outputs[-2], dim=1)                                   # pixel_acc (utils.py:29-35)
    valid = (labels >= 0).long()
    acc = torch.sum(valid * (preds == labels).long()).float() / (torch.sum(valid).float() + 1e-10)
    loss_s = sem_loss(outputs[:-1], labels, weight, cfg)
    loss_b = cfg['coeff_bce'] * weighted_bce(outputs[-1], bd_gt)
    filler = torch.ones_like(labels) * cfg['ignore_label']
    bd_label = torch.where(torch.sigmoid(outputs[-1][:, 0, :, :]) > cfg['bd_threshold'], labels, filler)
    loss_sb = sem_loss(outputs[-2], bd_label, weight, cfg)
    loss = loss_s + loss_b + loss_sb
    return torch.unsqueeze(loss, 0), outputs[:-1], acc, [loss_s, loss_b]


def synthetic_batch(n, ncls, h, w, seed, lowres_div=8, scale=3.0, aligned=False):
    """Seeded logits / labels (with a 255 ignore band) / boundary targets, as SURVEY.md section 8d config 5.
    aligned=True: blocky labels and logits that mostly predict them (a confident, late-training regime in
    which the OHEM k-th order statistic exceeds the 0.9 threshold)."""
    g = torch.Generator().manual_seed(seed)
    hl, wl = h // lowres_div, w // lowres_div
    xp = scale * torch.randn(n, ncls, hl, wl, generator=g)
    xm = scale * torch.randn(n, ncls, hl, wl, generator=g)
    xd = scale * torch.randn(n, 1, hl, wl, generator=g)
    labels = torch.randint(0, ncls, (n, h, w), generator=g)
    if aligned:
        low = torch.randint(0, ncls, (n, hl // 4, wl // 4), generator=g)
        low = low.repeat_interleave(4, 1).repeat_interleave(4, 2)
        labels = low.repeat_interleave(lowres_div, 1).repeat_interleave(lowres_div, 2).contiguous()
        onehot = F.one_hot(low, ncls).permute(0, 3, 1, 2).float()
        xm = 12.0 * onehot + torch.randn(n, ncls, hl, wl, generator=g)
        xp = 6.0 * onehot + torch.randn(n, ncls, hl, wl, generator=g)
    labels[:, : max(1, h // 16), :] = 255
    labels[:, :, -max(1, w // 32):] = 255
    bd = (torch.rand(n, h, w, generator=g) > 0.9).float()
    return [xp, xm, xd], labels, bd
