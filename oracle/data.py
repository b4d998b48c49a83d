"""Oracle: the reference's rigid-transform augmentation (data side).  Test infrastructure only.

All citations: ``Point_Cloud_Resistration/data_utils/Data_set_maker.py``.  Pinned by ``tests/golden/rigid_transform.npz``
(the unmodified reference run on seeded inputs, ``tests/golden/make_golden_data.py``).
"""
import numpy as np
import torch
import torch.nn.functional as F


def qmul(q, r):
    """Hamilton product as the reference writes it (:25-38: terms = r (x) q)."""
    t = torch.bmm(r.view(-1, 4, 1), q.view(-1, 1, 4))
    w = t[:, 0, 0] - t[:, 1, 1] - t[:, 2, 2] - t[:, 3, 3]
    x = t[:, 0, 1] + t[:, 1, 0] - t[:, 2, 3] + t[:, 3, 2]
    y = t[:, 0, 2] + t[:, 1, 3] + t[:, 2, 0] - t[:, 3, 1]
    z = t[:, 0, 3] - t[:, 1, 2] + t[:, 2, 1] + t[:, 3, 0]
    return torch.stack((w, x, y, z), dim=1).view(q.shape)


def euler_to_quaternion(e, order="xyz"):
    """:58-100 -- product of the axis quaternions in ``order``; sign flipped for the cyclic orders (:97-98)."""
    e = np.asarray(e, dtype=np.float64).reshape(-1, 3)
    half = {"x": e[:, 0] / 2, "y": e[:, 1] / 2, "z": e[:, 2] / 2}
    axis = {"x": 1, "y": 2, "z": 3}
    result = None
    for c in order:
        r = np.zeros((e.shape[0], 4))
        r[:, 0] = np.cos(half[c])
        r[:, axis[c]] = np.sin(half[c])
        result = r if result is None else qmul(torch.from_numpy(result).contiguous(), torch.from_numpy(r).contiguous()).numpy()
    if order in ("xyz", "yzx", "zxy"):
        result = result * -1
    return result


def create_random_transform(rng, max_rotation_deg=45, max_translation=1):
    """:154-171 with an explicit ``numpy.random.RandomState`` (the reference uses the global numpy RNG): Euler angles
    U(-max, max), translation sqrt(max_translation) * unit(U(-1,1)^3); returns the (1,7) float32 pose vector."""
    mr = np.pi / 180 * max_rotation_deg
    rot = rng.uniform(-mr, mr, [1, 3])
    trans = rng.uniform(-1, 1, [1, 3])
    trans = np.sqrt(max_translation) * (trans / np.linalg.norm(trans))
    return torch.tensor(np.concatenate([euler_to_quaternion(rot, "xyz"), trans], axis=1), dtype=torch.float32)


def qrot(q, v):
    """:40-52 -- v + 2 (w (q x v) + q x (q x v))."""
    qvec = q[..., 1:]
    uv = torch.cross(qvec, v, dim=-1)
    uuv = torch.cross(qvec, uv, dim=-1)
    return v + 2 * (q[..., :1] * uv + uuv)


def rigid_transform(source, poses):
    """``Dataset_Transformation.__call__`` (:221-230) batched: source (B,N,3), poses (B,7) ->
    (transformed (B,N,3), igt_rotation (B,3,3), igt_translation (B,1,3))."""
    quat = F.normalize(poses[:, 0:4], dim=1)  # create_pose_7d :173-181
    trans = poses[:, 4:]
    B, N, _ = source.shape
    out = qrot(quat.unsqueeze(1).expand(-1, N, -1), source) + trans.unsqueeze(1)
    eye = torch.eye(3, dtype=source.dtype).unsqueeze(0).expand(B, -1, -1)
    rot = qrot(quat.unsqueeze(1).expand(-1, 3, -1), eye).permute(0, 2, 1)  # quaternion_rotate(eye(3), igt).permute(1, 0)
    return out, rot, trans.unsqueeze(1)
