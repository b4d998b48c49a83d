"""Golden fixture for the data side: the reference's rigid-transform augmentation
(``Point_Cloud_Resistration/data_utils/Data_set_maker.py:137-230``) run UNMODIFIED in the build container on seeded inputs.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_data.py

``torch_geometric`` (dataset download / mesh sampling, out of scope) is not installed: it is stubbed with empty modules so
that ``Data_set_maker`` imports; only ``Dataset_Transformation`` and the quaternion helpers are executed."""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference/Point_Cloud_Resistration"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
for name in ("torch_geometric", "torch_geometric.transforms", "torch_geometric.datasets"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["torch_geometric.datasets"].ModelNet = object
sys.path.insert(0, os.path.join(REF, "data_utils"))
import Data_set_maker as dm  # noqa: E402

np.random.seed(1234)
torch.manual_seed(1234)
B, N = 6, 100
tf = dm.Dataset_Transformation(B, angle_range=45, translation_range=1)
src = torch.randn(B, N, 3)
poses = torch.cat(tf.transformations, 0)  # (B,7) un-normalised quaternion + translation
out, rot, trans = [], [], []
for i in range(B):
    tf.index = i
    out.append(tf(src[i]))
    rot.append(tf.igt_rotation)
    trans.append(tf.igt_translation)
np.savez(os.path.join(HERE, "rigid_transform.npz"), src=src.numpy(), poses=poses.numpy(), out=torch.stack(out).numpy(),
         igt_rotation=torch.stack(rot).numpy(), igt_translation=torch.stack(trans).numpy(),
         euler_in=np.array([[0.3, -0.2, 0.5], [-0.7, 0.1, 0.4]]),
         euler_quat=dm.euler_to_quaternion(np.array([[0.3, -0.2, 0.5], [-0.7, 0.1, 0.4]]), "xyz"))
print("wrote rigid_transform.npz", poses.shape)
