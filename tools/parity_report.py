#!/usr/bin/env python
"""Print per-stream and gradient errors of the CUDA path against the reference's FP64 golden run (fixture `ldc` etc.),
for the FFMA path (TDE_NO_TC=1) and the tcgen05 path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_de_b200 as T
from tests.golden_util import load, rel_l2, max_rel
from tests.test_jet_parity_gpu import product_workload, load_net

for fixture in sys.argv[1:] or ['ldc', 'allen_cahn']:
    for no_tc in ('1', '0'):
        os.environ['TDE_NO_TC'] = no_tc
        z, owl, params, groups = load(fixture)
        wl = product_workload(fixture)
        net, lins = load_net(wl, params, 'cuda')
        pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
        batch = T.make_batch(wl.input_vars, {g: x.cuda() for g, x in groups.items()})
        out = pinn.calculate(batch)
        loss = wl.build_losses(list(groups.keys()))(batch, out)
        total = wl.combine(loss.grouped_losses('loss_type'))
        total.backward()
        worst = max(((max_rel(out['all'][k].detach().cpu(), z['s64_' + k]), rel_l2(out['all'][k].detach().cpu(), z['s64_' + k]), k)
                     for k in pinn.derivatives))
        ref32 = max(((max_rel(z['s32_' + k], z['s64_' + k]), k) for k in pinn.derivatives))
        flat = torch.cat([p.grad.flatten().cpu() for lin in lins for p in (lin.weight, lin.bias)]).double()
        msg = f"{fixture:10s} {'FFMA ' if no_tc == '1' else 'tcgen05'}: worst stream max_rel {worst[0]:.2e} relL2 {worst[1]:.2e} ({worst[2]}); " \
              f"reference fp32 worst {ref32[0]:.2e} ({ref32[1]}); total rel err {abs(float(total) - float(z['total64'])) / abs(float(z['total64'])):.2e}"
        if 'gflat64_norm' in z:
            msg += f"; |grad| rel err {abs(float(flat.norm()) - float(z['gflat64_norm'])) / float(z['gflat64_norm']):.2e}"
        print(msg)
