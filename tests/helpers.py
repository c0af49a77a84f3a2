"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sub(t: torch.Tensor, n: int = 16384) -> np.ndarray:
    """Same deterministic strided subsample as tools/make_golden.py."""
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].to(torch.float32).cpu().numpy().copy()


def rel_err(a, b) -> float:
    """max|a-b| / max|b| -- the 'max relative error' of the north star (b = reference)."""
    a = torch.as_tensor(np.asarray(a, dtype=np.float64)) if not torch.is_tensor(a) else a.double().cpu()
    b = torch.as_tensor(np.asarray(b, dtype=np.float64)) if not torch.is_tensor(b) else b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load_golden(name: str):
    return np.load(os.path.join(GOLD, name))
