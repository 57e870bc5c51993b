"""TEST INFRASTRUCTURE ONLY - mints tests/golden/mpc_reference_qp.npz from the UNMODIFIED reference MPC set-up.

For a handful of problems (all contact multiplicities, saturating demands) and a set of force vectors each - random ones,
a strictly interior one, and ones that break exactly one constraint (fz just above 150, |fx| or |fy| just above 0.6 fz,
a non-zero swing force) next to their just-feasible twins - it records what /root/reference/misc/force_controller.py's own
constructor code evaluates (oracle/mpc_ref_shim.py): the objective value and whether every subject_to holds.

    python -m oracle.gen_golden_mpc
"""
import os

import numpy as np

from oracle import mpc_numpy as mpc
from oracle import mpc_ref_shim as shim
from tests import mpc_cases

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mpc_reference_qp.npz")
CONTACTS = ([1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 1], [1, 0, 0, 0], [0, 0, 0, 0], [1, 1, 0, 1])


def force_set(rng, x, ref, p, contact):
    contact = np.asarray(contact, float)
    stance = [l for l in range(4) if contact[l] == 1]
    inner = np.zeros((12, 5))
    for l in stance:
        inner[3 * l + 2] = 10.0
    opt = mpc.solve(x, ref, p, contact)
    base = 0.5 * opt + 0.5 * inner  # strictly inside the feasible set (or zero when every leg swings)
    out = [30.0 * rng.standard_normal((12, 5)), base]
    for l in stance[:2]:
        for frac, comp in ((0.61, 0), (0.59, 0), (-0.61, 1), (-0.59, 1)):
            f = base.copy()
            f[3 * l + comp, 1] = frac * f[3 * l + 2, 1]
            out.append(f)
        for fz in (150.5, 149.5, -0.5):
            f = base.copy()
            f[3 * l + 2, 3] = fz
            f[3 * l, 3] = f[3 * l + 1, 3] = 0.0
            out.append(f)
    for l in [l for l in range(4) if contact[l] == 0][:1]:
        f = base.copy()
        f[3 * l + 2, 0] = 1e-3  # a swing leg must carry no force
        out.append(f)
    return out


def main():
    rng = np.random.default_rng(20231017)
    rec = {k: [] for k in ("x", "body_ref", "p", "contact", "forces", "cost", "feasible")}
    for k, contact in enumerate(CONTACTS):
        x, ref, p = mpc_cases.problem(rng, lateral=[0.0, 0.5, 3.0][k % 3], height_error=k == 3)
        for f in force_set(rng, x, ref, p, contact):
            cost, feasible = shim.evaluate(f, x, ref, p, contact)
            for key, val in (("x", x), ("body_ref", ref), ("p", p), ("contact", np.asarray(contact, float)), ("forces", f), ("cost", cost),
                             ("feasible", feasible)):
                rec[key].append(val)
    np.savez_compressed(OUT, **{k: np.array(v) for k, v in rec.items()})
    print(f"{OUT}: {len(rec['cost'])} evaluations of the reference set-up, {int(np.sum(rec['feasible']))} feasible")


if __name__ == "__main__":
    main()
