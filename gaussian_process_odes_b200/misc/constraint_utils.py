"""Positive-parameter constraint (mirror of reference ``src/misc/constraint_utils.py:5-13``)."""
import torch
import torch.nn.functional as F

LOWER = 1e-12


def softplus(x):
    return F.softplus(x) + LOWER


def invsoftplus(x):
    eps = torch.tensor(torch.finfo(x.dtype).eps).to(x)
    shifted = torch.max(x - LOWER, eps)
    return shifted + torch.log(-torch.expm1(-shifted))
