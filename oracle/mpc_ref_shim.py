"""TEST INFRASTRUCTURE ONLY - evaluates the UNMODIFIED reference force MPC set-up numerically.

/root/reference/misc/force_controller.py::StanceController builds its QP symbolically with CasADi (`from casadi import *`,
:12) and hands it to qpOASES; neither is installed offline.  Its constructor, however, only ever *composes* expressions:
Opti.variable / Opti.parameter placeholders, vertcat / horzcat / mtimes / transpose / inv / skew / cos / sin / if_else, and
Opti.bounded / subject_to / minimize.  This module plants a `casadi` stand-in whose placeholders already carry NUMBERS
(numpy.matrix, which slices like a CasADi matrix: always 2-D) and whose operators are the NumPy ones.  Running the
reference's own `StanceController.__init__` under it therefore evaluates, with the reference's own code and operand order,

    * the objective value `cost` for the given forces and parameters          (force_controller.py:65-104), and
    * whether the given forces satisfy every `subject_to` constraint          (force_controller.py:106-156).

That pins the oracle's restatement of the QP (oracle/mpc_numpy.py: rollout_cost, build_qp, constraints) to the reference
itself; what stays unpinned is only the solver (qpOASES), and the QP is strictly convex, so its minimiser is unique.
Nothing from the reference is copied; the module is loaded from where it lies.  Used by oracle/gen_golden_mpc.py (mints
tests/golden/mpc_reference_qp.npz) and by the CPU tests when the reference tree is present.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("OPTISTATE_REF", "/root/reference")
_PATH = os.path.join(REF_ROOT, "misc", "force_controller.py")


def available() -> bool:
    return os.path.isfile(_PATH)


def _m(a):
    return np.matrix(np.atleast_2d(np.asarray(a, dtype=float)))


class _NumericOpti:
    """Opti stand-in: placeholders come from the value queues filled by evaluate()."""
    variables: list = []
    parameters: list = []

    def __init__(self, kind):
        assert kind == "conic"
        self.constraints, self.cost = [], None

    def variable(self, n, m):
        v = _NumericOpti.variables.pop(0)
        assert v.shape == (n, m), (v.shape, n, m)
        return _m(v)

    def parameter(self, n, m):
        v = _NumericOpti.parameters.pop(0)
        if (n, m) == (1, 1):
            return float(np.asarray(v).reshape(-1)[0])  # a CasADi 1x1 acts as a scalar in `A * dt`
        assert np.shape(v) == (n, m), (np.shape(v), n, m)
        return _m(v)

    def set_value(self, *a):
        pass

    def bounded(self, lo, x, hi):
        return np.logical_and(np.asarray(lo) <= np.asarray(x), np.asarray(x) <= np.asarray(hi))

    def subject_to(self, cond):
        self.constraints.append(bool(np.all(np.asarray(cond))))

    def minimize(self, cost):
        self.cost = float(np.asarray(cost).reshape(-1)[0])

    def solver(self, *a, **k):
        pass


def _skew(v):
    a, b, c = (float(t) for t in np.asarray(v, dtype=float).reshape(-1))
    return _m([[0.0, -c, b], [c, 0.0, -a], [-b, a, 0.0]])


def _stub():
    cas = types.ModuleType("casadi")
    inner = types.SimpleNamespace(Opti=_NumericOpti)
    names = {
        "casadi": inner,
        "vertcat": lambda *a: _m(np.vstack([np.atleast_2d(np.asarray(x, dtype=float)) for x in a])),
        "horzcat": lambda *a: _m(np.hstack([np.atleast_2d(np.asarray(x, dtype=float)) for x in a])),
        "mtimes": lambda a, b: _m(np.asarray(a, dtype=float) @ np.asarray(b, dtype=float)),
        "if_else": lambda c, a, b: a if bool(np.all(np.asarray(c))) else b,
        "cos": np.cos, "sin": np.sin, "tan": np.tan,
        "transpose": lambda a: _m(np.asarray(a, dtype=float).T),
        "inv": lambda a: _m(np.linalg.inv(np.asarray(a, dtype=float))),
        "skew": _skew,
    }
    for k, v in names.items():
        setattr(cas, k, v)
    cas.__all__ = list(names)
    return cas


def _load_controller():
    """A private copy of the reference module bound to the numeric stand-in (sys.modules['casadi'] is restored)."""
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    saved = sys.modules.get("casadi")
    sys.modules["casadi"] = _stub()
    try:
        spec = importlib.util.spec_from_file_location("_ref_force_controller_numeric", _PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("casadi", None)
        else:
            sys.modules["casadi"] = saved
    return mod


def evaluate(forces, x, body_ref, p, contact, dt=0.01):
    """(cost, feasible) of the forces (12, N) under the reference's own set-up code, with the parameters filled the way
    Kalman_Filter.predict_mpc fills them (kalman_filter.py:64-77 weights, :141-146 horizon arrays)."""
    import warnings

    warnings.filterwarnings("ignore", category=PendingDeprecationWarning)  # numpy.matrix is what slices like a CasADi matrix
    mod = _load_controller()
    N = 5
    Q = np.diag([10.0, 10.0, 10.0, 100.0, 100.0, 100.0, 1.0, 1.0, 5.0, 1.0, 1.0, 1.0])  # kalman_filter.py:64-65
    R = np.diag([0.000001] * 12)                                                          # kalman_filter.py:66-70
    forces = np.asarray(forces, float).reshape(12, N)
    body = np.concatenate([np.asarray(x, float).reshape(12, 1), np.asarray(body_ref, float).reshape(12, N)], axis=1)  # :143-144
    feet = np.repeat(np.asarray(p, float).reshape(12, 1), N + 1, axis=1)                                              # :141-142
    cont = np.repeat(np.asarray(contact, float).reshape(4, 1), N, axis=1)                                             # :145-146
    _NumericOpti.variables = [forces[0:3].copy(), forces[3:6].copy(), forces[6:9].copy(), forces[9:12].copy()]        # f1..f4
    _NumericOpti.parameters = [body, feet, cont, np.array([[dt]])]                           # body_mpc, p_mpc, contact_mpc, dt_param
    sc = mod.StanceController(N, Q, R, Q, dt)                                                # P = Q, kalman_filter.py:71
    assert not _NumericOpti.variables and not _NumericOpti.parameters
    return sc.opti.cost, all(sc.opti.constraints)
