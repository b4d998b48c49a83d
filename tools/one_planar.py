"""A handful of Planar sphere-map calls for an ncu capture (csrc/planar.cu): clouds layout 32 x 1024 (one CTA per cloud), one cloud
of 262144 points (8-CTA cluster), points layout 2^20 x 3, three layers, forward + backward each.
    ncu --set full --clock-control none --import-source on -k regex:planar_ -c 9 -f -o gpurun_out/planar python tools/one_planar.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
phi = shwd.losses.Norm_Flow_structure(flow_name="Planar", n_flow_layer=3).to(dev)
for shape in ((32, 1024, 3), (1, 262144, 3), (1 << 20, 3)):
    x = torch.randn(*shape, device=dev, requires_grad=True)
    y = phi(x)
    torch.autograd.grad(y.square().sum(), [x] + list(phi.parameters()))
torch.cuda.synchronize()
print("ok")
