"""Device-resident data side for registration training.

The reference builds every batch on the CPU: ``Dataset_pytorch.__getitem__`` applies a per-item random rigid transform
(``Dataset_Transformation``, data_utils/Data_set_maker.py:137-230) inside two DataLoader workers and the batch is then
copied to the GPU each step (train_W_COS.py:161-164, 363-364).  Here the clouds live on the device and one kernel launch
(``csrc/rigid.cu``) transforms a whole batch; the poses are drawn exactly like ``create_random_transform`` (:154-171).
"""
import numpy as np
import torch

from . import _lib
from .ops import _ptr, _stream


def euler_to_quaternion_xyz(e):
    """(n,3) Euler angles -> (n,4) quaternions (w,x,y,z) as ``euler_to_quaternion(e, "xyz")`` (:58-100): q = -(qx*qy*qz)
    in the reference's ``qmul`` convention (terms = r (x) q)."""
    e = np.asarray(e, dtype=np.float64).reshape(-1, 3)

    def qmul(q, r):
        w = r[:, 0] * q[:, 0] - r[:, 1] * q[:, 1] - r[:, 2] * q[:, 2] - r[:, 3] * q[:, 3]
        x = r[:, 0] * q[:, 1] + r[:, 1] * q[:, 0] - r[:, 2] * q[:, 3] + r[:, 3] * q[:, 2]
        y = r[:, 0] * q[:, 2] + r[:, 1] * q[:, 3] + r[:, 2] * q[:, 0] - r[:, 3] * q[:, 1]
        z = r[:, 0] * q[:, 3] - r[:, 1] * q[:, 2] + r[:, 2] * q[:, 1] + r[:, 3] * q[:, 0]
        return np.stack((w, x, y, z), axis=1)

    out = None
    for k in range(3):
        r = np.zeros((e.shape[0], 4))
        r[:, 0] = np.cos(e[:, k] / 2)
        r[:, 1 + k] = np.sin(e[:, k] / 2)
        out = r if out is None else qmul(out, r)
    return -out


def random_poses(count, angle_range=45, translation_range=1, rng=None):
    """(count,7) float32 poses drawn like ``Dataset_Transformation.create_random_transform`` (:154-171), one by one from
    ``rng`` (a ``numpy.random.RandomState``; the reference uses numpy's global stream)."""
    rng = rng if rng is not None else np.random
    mr = np.pi / 180 * angle_range
    vecs = []
    for _ in range(count):
        rot = rng.uniform(-mr, mr, [1, 3])
        trans = rng.uniform(-1, 1, [1, 3])
        trans = np.sqrt(translation_range) * (trans / np.linalg.norm(trans))
        vecs.append(np.concatenate([euler_to_quaternion_xyz(rot), trans], axis=1))
    return torch.tensor(np.concatenate(vecs, 0), dtype=torch.float32)


def rigid_transform(source, poses, noise_std=0.0, seed=0):
    """source (B,N,3) CUDA, poses (B,7) -> (transformed (B,N,3), igt_rotation (B,3,3), igt_translation (B,1,3)):
    ``Dataset_Transformation.__call__`` (:221-230) for the whole batch in one launch."""
    if not source.is_cuda:
        raise RuntimeError("source must live on a CUDA device: no CPU fallback")
    src = source.contiguous().float()
    pose = poses.to(device=src.device, dtype=torch.float32).contiguous()
    B, N, _ = src.shape
    out = torch.empty_like(src)
    rot = torch.empty(B, 3, 3, device=src.device, dtype=torch.float32)
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().shwd_rigid_transform(_ptr(src), _ptr(pose), B, N, float(noise_std), int(seed) & (2 ** 64 - 1), _ptr(out),
                                                   _ptr(rot), _stream()), "shwd_rigid_transform")
    return out, rot, pose[:, 4:].unsqueeze(1)


class DeviceRegistrationPairs:
    """Drop-in for ``Dataset_pytorch`` + ``DataLoader`` (:238-259, train_W_COS.py:357-364) with the clouds resident on the
    device: ``batch(indices)`` returns the tuple a collated ``__getitem__`` gives -- (target, transformed_source,
    igt_rotation, igt_translation) -- without leaving the GPU.  Each item keeps its own fixed pose, as in the reference."""

    def __init__(self, sources, targets, angle_range=45, translation_range=1, noise_sigma=0.0, seed=1234):
        self.sources = sources.contiguous().float()
        self.targets = targets.contiguous().float()
        self.poses = random_poses(self.sources.shape[0], angle_range, translation_range, np.random.RandomState(seed)).to(self.sources.device)
        self.noise_sigma, self.seed = float(noise_sigma), int(seed)

    def __len__(self):
        return self.sources.shape[0]

    def batch(self, indices):
        idx = torch.as_tensor(indices, device=self.sources.device, dtype=torch.long)
        out, rot, trans = rigid_transform(self.sources[idx], self.poses[idx], self.noise_sigma, self.seed)
        return self.targets[idx], out, rot, trans

    def batches(self, batch_size, shuffle=True, generator=None):
        n = len(self)
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        for i in range(0, n, batch_size):
            yield self.batch(order[i:i + batch_size])
