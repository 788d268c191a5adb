import numpy as np
import torch

from oracle import tt_oracle as orc


def to_np_cores(cores):
    return [c.detach().cpu().numpy() for c in cores]


def relerr64(X, cores):
    """fp64 relative reconstruction error on the host (oracle definition)."""
    return orc.relative_error(np.asarray(X), to_np_cores(cores))


def ranks_of(cores):
    return [1] + [int(c.shape[2]) for c in cores]
