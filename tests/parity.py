"""Parity metrics shared by the oracle and CUDA tests (SURVEY.md section 8(c) tolerances).

States:      max_t |x - x_ref| / max_t |x_ref|   per state, then max over states.
Covariance:  max |P - P_ref| / max |P_ref|  (the headline bound) and the stricter
             max |P_ij - Pref_ij| / sqrt(Pref_ii * Pref_jj)  (error relative to the entry's own scale).
Element-wise relative error on near-zero entries is deliberately not used.
"""
import numpy as np

FP64_TOL = 1e-9  # BASELINE.json north_star: FP64 within 1e-9 relative on states and covariances
# FP32 over 10k steps (SURVEY 8(c) probe: 3.4e-6 / 2.5e-5 / 1.0e-4 observed in emulation)
FP32_TOL_X, FP32_TOL_P, FP32_TOL_TRACE = 2e-5, 2e-4, 5e-4


def state_err(x, x_ref, x_absmax=None):
    """x, x_ref: [T, 12]."""
    scale = np.abs(x_ref).max(axis=0) if x_absmax is None else np.asarray(x_absmax)
    scale = np.where(scale > 0, scale, 1.0)
    return float((np.abs(x - x_ref).max(axis=0) / scale).max())


def cov_err(P, P_ref):
    P, P_ref = np.asarray(P).reshape(-1, 12, 12), np.asarray(P_ref).reshape(-1, 12, 12)
    e_max = max(float(np.abs(a - b).max() / np.abs(b).max()) for a, b in zip(P, P_ref))
    e_corr = 0.0
    for a, b in zip(P, P_ref):
        d = np.sqrt(np.abs(np.diag(b)))
        e_corr = max(e_corr, float((np.abs(a - b) / np.outer(d, d)).max()))
    return e_max, e_corr


def rel_err(a, a_ref):
    a, a_ref = np.asarray(a, float), np.asarray(a_ref, float)
    return float(np.abs(a - a_ref).max() / np.abs(a_ref).max())
