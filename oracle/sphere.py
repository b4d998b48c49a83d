"""Oracle: sphere map (centre + project on the unit sphere) and the flow regulariser.  Test infrastructure only."""
import torch


def sphere_map(x, center=True, normalize=True, eps=1e-8):
    """Centroid subtraction then hard normalisation.

    Centring follows ``Point_Cloud_Resistration/train_W_COS.py:167-168`` (``x - mean(x, dim=1, keepdim=True)``).
    Normalisation follows what ``F.cosine_similarity`` does internally at
    ``Point_Cloud_Resistration/losses/s2_wasserstein.py:122``: ``x / max(||x||_2, 1e-8)`` (torch 2.11 normalises
    each operand first, then takes the dot product -- SURVEY.md B.1).
    Accepts (B,N,3) or (N,3).
    """
    if center:
        x = x - torch.mean(x, dim=-2, keepdim=True)
    if normalize:
        x = x / torch.linalg.vector_norm(x, dim=-1, keepdim=True).clamp_min(eps)
    return x


def flow_regularization(x):
    """``sum_{b,n} | ||x_bn||_2 - 1 |`` -- ``Point_Cloud_Resistration/losses/s2_wasserstein.py:224-232``."""
    if x.dim() == 2:
        x = x.unsqueeze(0)
    return torch.sum(torch.abs(torch.linalg.vector_norm(x, dim=-1) - 1))
