"""The strided sampling oracle/make_golden.py uses for big tensors (fixtures hold samples + norms)."""
import numpy as np


def grad_sample(t, n=64):
    f = np.asarray(t.detach().cpu().double().reshape(-1)) if hasattr(t, "detach") else np.asarray(t, np.float64).reshape(-1)
    step = max(1, f.size // n)
    return f[::step][:n].copy()
