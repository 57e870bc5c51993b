"""Seeded force-MPC problems shared by the CPU (oracle) and GPU tests of SURVEY 8(f) row 3."""
import numpy as np

NOMINAL_FEET = np.array([0.2, 0.15, -0.28, 0.2, -0.15, -0.28, -0.2, 0.15, -0.28, -0.2, -0.15, -0.28])


def problem(rng, lateral=0.0, height_error=False, ref_on_ground=False):
    """One problem: current state x (12), horizon reference body_ref (12, 5), body-frame feet p (12)."""
    x = np.array([0.02, -0.03, 0.1, 0, 0, 0.27, 0.1, -0.1, 0.05, 0.2, -0.1, 0.0]) + 0.01 * rng.standard_normal(12)
    ref = np.zeros((12, 5))
    ref[5] = 0.28
    ref[0:3] = 0.02 * rng.standard_normal((3, 5))
    ref[9] = lateral                                   # demanded forward velocity: large values saturate friction and fz
    ref[3] = x[3] + lateral * 0.01 * np.arange(1, 6)
    if height_error:
        x[5] = 0.1
    if ref_on_ground:
        ref[5] = 0.0
    p = NOMINAL_FEET + 0.02 * rng.standard_normal(12)
    return x, ref, p


def batch(n, seed=0):
    """n problems cycling through all 16 contact patterns and the demand / error variants; arrays in device layout
    x [12, n], body_ref [5, 12, n], p [12, n], contact [4, n]."""
    rng = np.random.default_rng(seed)
    xs, refs, ps, cs = [], [], [], []
    for k in range(n):
        contact = np.array([(k >> b) & 1 for b in range(4)], float)
        x, ref, p = problem(rng, lateral=[0.0, 0.5, 3.0, -2.0][(k // 16) % 4], height_error=k % 7 == 0, ref_on_ground=k % 11 == 0)
        xs.append(x), refs.append(ref.T), ps.append(p), cs.append(contact)
    return (np.stack(xs, axis=1), np.stack(refs, axis=2), np.stack(ps, axis=1), np.stack(cs, axis=1))
