import torch, sys, os
sys.path.insert(0, '/root/repo')
import torch_de_b200 as tde
from torch_de_b200.networks import MLP
torch.manual_seed(0)
net = MLP(2, 1, 256, 8).cuda()
names = sys.argv[1].split(',')
eng = tde.DE_Getter(net)
eng.set_vars(['x', 't'], ['u'])
eng.set_derivatives(names)
x = torch.rand(200000, 2, device='cuda')
for i in range(3):
    out = eng.calculate(x)
    loss = sum((v**2).mean() for v in out.values() if torch.is_tensor(v))
    loss.backward()
torch.cuda.synchronize()
