import os, sys
sys.path.insert(0, "/root/repo")
import torch, torch.nn.functional as F
import shwd
dev = torch.device("cuda:0")
n = 4000
g = torch.Generator().manual_seed(11)
x = F.normalize(torch.randn(1, n, 3, generator=g), dim=-1).to(dev)
y = F.normalize(torch.randn(1, n, 3, generator=g) + 0.2, dim=-1).to(dev)
U, _ = torch.linalg.qr(torch.randn(8, 3, 2, generator=g)); U = U.to(dev)
ku = shwd.ops.ProjectCircleFn.apply(x, U).reshape(8, n)
kv = shwd.ops.ProjectCircleFn.apply(y, U).reshape(8, n)
us, vs = torch.sort(ku, -1)[0].contiguous(), torch.sort(kv, -1)[0].contiguous()
shwd.ops.CircularWpFn.apply(us, vs, 2.0, -1.0, 1.0, 1e-7)
torch.cuda.synchronize()
