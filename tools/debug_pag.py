import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import pidnet_oracle as O
from tests.test_net_gpu import build
dev = torch.device('cuda:0')
for impl in (1, 0):
  for lanes in (1, 3):
    model, sd = build('tiny_s', 5, True, 11, dev, conv_impl=impl, lanes=lanes)
    x = torch.randn(2, 3, 64, 128, generator=torch.Generator().manual_seed(5))
    taps = {}
    with torch.no_grad():
        ref = O.pidnet_forward(sd, x, taps=taps)
        got = model(x.to(dev))
    torch.cuda.synchronize()
    e = {k: model.debug_tensor(k) for k in ['layer3_', 'layer3', 'pag3', 'layer2', 'layer1']}
    c = O._Ctx(sd)
    with torch.no_grad():
        y = O._bn(c, O._conv(c, e['layer3'], 'compression3.0'), 'compression3.1')
        loc = F.relu(O.pagfm(c, e['layer3_'], y, 'pag3'))
    print(f'impl={impl} lanes={lanes} pag3 vs oracle {O.rel_l2(e["pag3"], taps["pag3"]):.4g}  pag3 vs local-recompute {O.rel_l2(e["pag3"], loc):.4g}  layer1(relu) {O.rel_l2(e["layer1"], F.relu(taps["layer1"])):.4g}')
    d = (e['pag3'] - loc).abs()
    print('   max abs diff', d.max().item(), 'at', (d == d.max()).nonzero()[0].tolist(), 'frac>0.1:', (d > 0.1).float().mean().item())
