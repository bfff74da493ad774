"""Drop-in for ``from nt_xent import NT_Xent`` (DMGAN+CLIP/code/pretrain_DAMSM.py:35, trainer.py:32;
definition at nt_xent.py:4-35)."""
import importlib
import os
import sys

import torch
import torch.nn as nn

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("t2i_clip-gan_b200")


class NT_Xent(nn.Module):
    """Same constructor and call as the reference (nt_xent.py:6-16): ``NT_Xent(batch_size, temperature, mask,
    device)(z_i, z_j) -> 0-d loss``.  ``mask`` must be the one ``masks.mask_correlated_samples(_2)`` builds (diagonal
    and +-batch_size diagonals removed) -- the kernel hard-wires it; any other mask raises ``ValueError``."""

    def __init__(self, batch_size, temperature, mask, device):
        super().__init__()
        self.batch_size = int(batch_size)
        self.temperature = temperature
        self.device = device
        if mask is not None:
            m = torch.as_tensor(mask).cpu()
            if m.shape != (2 * self.batch_size, 2 * self.batch_size) or \
                    not torch.equal(m.bool(), _pkg.standard_ntxent_mask(self.batch_size)):
                raise ValueError("NT_Xent: only the mask of masks.mask_correlated_samples is supported")
        self.mask = mask

    def forward(self, z_i, z_j):
        if z_i.shape[0] != self.batch_size:
            raise ValueError(f"NT_Xent: batch {z_i.shape[0]} does not match batch_size={self.batch_size}")
        return _pkg.nt_xent(z_i, z_j, self.temperature)
