"""Minimal stand-in so ``from pytorch3d.loss import chamfer_distance`` (train_CD.py:13, main_rotation.py:15) resolves
to the B200 Chamfer kernel.  Only use it (``PYTHONPATH=<repo>/dropin_pytorch3d``) where the real pytorch3d is absent."""
