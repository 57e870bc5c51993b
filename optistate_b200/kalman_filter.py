"""Drop-in for the reference's `Kalman_Filter` (/root/reference/kalman_filter/kalman_filter.py:7-202).

Same constructor (no arguments, reads settings.INITIAL_PARAMS), same attributes (NumPy arrays the caller may
read and overwrite between calls: x, z, H, P, Q, R, P_trace, F, B, g, dt, m, inertia_rot, x_model, K, K_gain,
F_d, B_d), same methods and the same side effects:
  * `predict` rotates the caller's `p` into the world frame in place (force_controller.py:274-277);
  * `set_measurements` writes into `self.z` in place (kalman_filter.py:113-117);
  * `predict` / `update` rebind `self.x` / `self.P` to new arrays (kalman_filter.py:133-135,169-172), so the
    STARTING_STATE / Q aliases are only mutated by callers that assign through `KF.x[:] = ...`.
Every numerical method runs on the GPU through the C ABI (one trajectory, one step, the JOINT kernel which keeps
the reference's operand order on a full non-symmetrised P); there is no CPU implementation behind it.  A method is one call of
the binding on a pinned block the kernels write their results into, and `predict` already launches the update that follows
it (see `_buffers`): 11 - 12 k steps/s stepped the way the reference driver steps it, against 4.7 - 4.9 k for the reference
class on one core of the same box.  For many trajectories or many steps use `optistate_b200.kf_batch`, which runs the
same arithmetic in one launch.

Errors follow the reference: `numpy.linalg.LinAlgError` when S is not invertible (kalman_filter.py:168) and
`ValueError` when no foot is in stance (kalman_filter.py:97-103).

The forces inside `predict_mpc` come from the reference's convex MPC (CasADi + qpOASES, force_controller.py:15-225).  Here
they are solved for on the device by optistate_b200.mpc.mpc_forces (same QP; dual active-set kernels with an interior
point behind them; parity with qpOASES unpinned, see there), unless forces are passed explicitly (`f=`) or an injectable `force_provider` is set.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _native as nv
from .settings import INITIAL_PARAMS

_SEL = [0, 1, 2, 5, 6, 7, 8, 9, 10, 11]
_F8 = np.dtype(np.float64)


class Kalman_Filter:
    def __init__(self, force_provider=None, device=None):
        self.x = INITIAL_PARAMS.STARTING_STATE  # aliased on purpose, kalman_filter.py:10
        self.z = np.zeros((10, 1))
        self.H = np.zeros((10, 12), dtype=np.int64)
        self.H[np.arange(10), _SEL] = 1
        self.P = INITIAL_PARAMS.P
        self.Q = INITIAL_PARAMS.Q
        self.R = INITIAL_PARAMS.R
        self.P_trace = np.trace(self.P)
        self.m = INITIAL_PARAMS.ROBOT_MASS
        self.inertia_rot = INITIAL_PARAMS.INERTIA_ROT
        self.identity_large = np.eye(12, 12)
        self.F = np.zeros((12, 12))
        self.F[3:6, 9:12] = np.eye(3)
        self.B = np.zeros((12, 12))
        for leg in range(4):
            self.B[9:12, 3 * leg:3 * leg + 3] = np.eye(3) / self.m
        self.g = np.array([0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, INITIAL_PARAMS.GRAVITY]).reshape(12, 1)
        self.dt = INITIAL_PARAMS.DT_mpc
        self.x_model = INITIAL_PARAMS.STARTING_STATE
        self.K = np.zeros((12, 10))
        self.K_gain = 0.0
        self.F_d = self.identity_large + self.dt * self.F
        self.B_d = self.dt * self.B
        self.f = np.zeros((12, 1))
        # attributes of the reference's MPC set-up (kalman_filter.py:58-77).  The QP lives on the device (optistate_b200.mpc), so
        # there is no CasADi StanceController object behind `stance_controller`; the horizon buffers keep the reference's shapes
        # and are filled by predict_mpc as the reference fills them (kalman_filter.py:141-146).
        self.stance_controller = None
        self.p_mpc = np.zeros((12, 6))
        self.body_mpc = np.zeros((12, 6))
        self.contact_mpc = np.zeros((4, 5))
        self.zero_mat = np.zeros((3, 3))
        self.identity = np.eye(3)
        self.identity_m = np.eye(3) / self.m
        self.force_provider = force_provider
        self.status = 0
        self._device = device

    # ------------------------------------------------------------------ device plumbing
    def _dev(self):
        nv.require_cuda()
        return torch.device("cuda", torch.cuda.current_device()) if self._device is None else torch.device(self._device)

    @staticmethod
    def _col(a, n):
        return np.asarray(a, dtype=np.float64).reshape(n)

    # One pinned host block with a fixed layout (inputs, then the outputs of predict / update / get_odom; offsets shared with
    # csrc/torch_binding.cpp:kf_class_step) and one device block, allocated once per instance.  A method packs its inputs through
    # NumPy views of the block, makes ONE call of the binding (upload, launches, results, one stream synchronisation - all below the
    # Python dispatcher) and reads its outputs from views again.  `transfer` picks how the bytes cross the host interface:
    # "copy" = asynchronous copies both ways, "store" = the kernels write results and status straight into the pinned block
    # (zero-copy stores, no download), "mapped" = they read their inputs from it as well (no upload either).
    # predict / predict_mpc also run the update that normally follows them, on the prediction's outputs and the current z and R, in
    # the same call (two launches back to back, no host round trip between them); update() hands that result out when x, P, z and R
    # are still what that launch saw, and launches on its own otherwise - same values either way (same kernel, same inputs).
    _LAYOUT = {"x": (0, 12), "P": (12, 144), "Q": (156, 144), "R": (300, 100), "p": (400, 12), "f": (412, 12), "z": (424, 10),
               "body_ref": (434, 12), "imu": (446, 6), "dp": (452, 12), "contact": (464, 4),
               "pred_x": (468, 12), "pred_P": (480, 144), "pred_p_world": (624, 12), "pred_trace": (636, 1),
               "upd_x": (637, 12), "upd_P": (649, 144), "upd_K": (793, 120), "upd_trace": (913, 1), "upd_kgain": (914, 1), "odom": (915, 4)}
    _BLOCK = 920
    _MODES = {"copy": 0, "store": 1, "mapped": 2}
    transfer = "store"

    _v = None

    def _buffers(self):
        if self._v is None:  # first device call of this instance: the blocks live on the device that is current now (or `device`)
            dev = self._dev()
            self._h_block = torch.zeros(self._BLOCK, dtype=torch.float64).pin_memory()
            self._d_block = torch.zeros(self._BLOCK, dtype=torch.float64, device=dev)
            self._h_status = torch.zeros(4, dtype=torch.int32).pin_memory()
            self._d_status = torch.zeros(4, dtype=torch.int32, device=dev)
            blk = self._h_block.numpy()
            self._v = {k: blk[o:o + n] for k, (o, n) in self._LAYOUT.items()}
            shapes = {"P": (12, 12), "Q": (12, 12), "R": (10, 10)}
            self._vs = {k: v.reshape(shapes.get(k, (v.size, 1))) for k, v in self._v.items()}
            self._st = self._h_status.numpy()
            self._call = nv.ext().kf_class_step
            self._ahead = None

    def _put(self, name, a, n):
        v = self._vs[name]  # the view in the shape the reference's arrays have: no reshape on the common path
        if type(a) is np.ndarray and a.shape == v.shape:
            v[...] = a
        else:
            self._v[name][:] = np.asarray(a, dtype=np.float64).reshape(n)

    def _run(self, ops, cov_model=nv.COV_PREDICT):
        I, g = self.inertia_rot, self.g
        consts = [float(self.dt), float(self.m), float(I[0][0]), float(I[1][1]), float(I[2][2]), float(g[11][0] if np.ndim(g) == 2 else g[11])]
        nv.check(int(self._call(ops, cov_model, consts, self._h_block, self._d_block, self._h_status, self._d_status, self._MODES[self.transfer])),
                 "optistate_kf_measure" if ops == nv.PHASE_MEASURE else "optistate_kf_batch")

    def _put_state(self):
        self._put("x", self.x, 12)
        self._put("P", self.P, 144)
        self._put("Q", self.Q, 144)
        self._put("R", self.R, 100)

    # ------------------------------------------------------------------ reference surface
    def get_odom(self, p_cur, dp_cur, contact_cur, imu):
        """kalman_filter.py:79-105 -> (4,1) array [z, vx, vy, vz]."""
        self._buffers()
        self._put("imu", imu, 6)
        self._put("p", p_cur, 12)
        self._put("dp", dp_cur, 12)
        self._put("contact", contact_cur, 4)
        self._run(nv.PHASE_MEASURE)
        self.status = int(self._st[2])
        if self.status & nv.ST_ALL_SWING:
            raise ValueError("setting an array element with a sequence. The requested array has an inhomogeneous shape "
                             "(all four feet in swing; the reference fails here too, kalman_filter.py:97-103)")
        return self._v["odom"].copy().reshape(4, 1)

    def set_measurements(self, imu, odom):
        """kalman_filter.py:108-117 (in-place scatter into self.z)."""
        imu, odom = np.asarray(imu), np.asarray(odom)
        self.z[0:3] = imu[0:3].reshape(3, 1)
        self.z[4:7] = imu[3:6].reshape(3, 1)
        self.z[3] = odom[0]
        self.z[7:10] = odom[1:].reshape(3, 1)

    def _refresh_model_matrices(self, angles, exp_form):
        """The host-side attributes the reference refreshes in predict / predict_mpc (kalman_filter.py:125-128,153-157): F carries R^T,
        F_d and B_d are rebound to new arrays.  (The filter evaluates its own R on the device; this is bookkeeping for callers that read them.)"""
        a, b, c = float(angles[0]), float(angles[1]), float(angles[2])
        sa, ca, sb, cb, sc, cc = math.sin(a), math.cos(a), math.sin(b), math.cos(b), math.sin(c), math.cos(c)
        self.F[0:3, 6:9] = ((cc * cb, sc * cb, -sb),
                            (cc * (sb * sa) - sc * ca, sc * (sb * sa) + cc * ca, cb * sa),
                            (cc * (sb * ca) + sc * sa, sc * (sb * ca) - cc * sa, cb * ca))
        self.F_d = np.exp(self.dt * self.F) if exp_form else self.identity_large + self.dt * self.F
        self.B_d = self.dt * self.B

    def _predict_call(self, p, f, cov_model, body_ref=None):
        """The device part of predict / predict_mpc, with the update that follows it computed ahead (see _buffers)."""
        self._buffers()
        self._put_state()
        self._put("p", p, 12)
        self._put("f", f, 12)
        self._put("z", self.z, 10)
        if body_ref is not None:
            self._put("body_ref", body_ref, 12)
        self._ahead = None
        self._run(nv.PHASE_PREDICT | nv.PHASE_UPDATE, cov_model)
        v = self._v
        self.status = int(self._st[0])
        p[...] = v["pred_p_world"].reshape(np.shape(p))
        self.x = v["pred_x"].copy().reshape(12, 1)
        self.P = v["pred_P"].copy().reshape(12, 12)
        self.x_model = self.x.copy()
        # what update() must still find in place to hand out the result computed ahead: the state this call produced, and the z and R
        # the second launch read (compared as bytes: stricter than ==, and cheap)
        self._ahead = (self.x.tobytes(), self.P.tobytes(), v["z"].tobytes(), v["R"].tobytes())

    def predict(self, p, f):
        """kalman_filter.py:119-138; p (12,1) is rotated into the world frame in place."""
        prior = self._col(self.x, 12).copy()
        self._predict_call(p, f, nv.COV_PREDICT)
        self._refresh_model_matrices(prior[0:3], exp_form=False)
        self.P_trace = float(self._v["pred_trace"][0])

    def predict_mpc(self, p, body_ref, cur_contact, f=None):
        """kalman_filter.py:140-162 with the QP replaced by `force_provider` (or an explicit f):
        covariance by F_d = exp(dt F) element-wise with R from body_ref, mean by next_state with R from x."""
        if f is None:
            if self.force_provider is not None:
                f = self.force_provider(np.asarray(p), np.asarray(body_ref), np.asarray(cur_contact), np.asarray(self.x))
            else:  # kalman_filter.py:141-152: body_mpc = [x | body_ref], feet and contact held over the horizon
                from .mpc import HORIZON, mpc_forces

                self._dev()  # fails loudly without a GPU: there is no CPU solver
                ref = np.asarray(body_ref, float)
                ref = ref if ref.ndim == 2 and ref.shape == (12, HORIZON) else np.repeat(ref.reshape(12, 1), HORIZON, axis=1)
                forces, st = mpc_forces(np.asarray(self.x, float).reshape(12, 1), ref.T.reshape(HORIZON, 12, 1),
                                        np.asarray(p, float).reshape(12, 1), np.asarray(cur_contact, float).reshape(4, 1),
                                        dt=float(self.dt), mass=float(self.m), inertia=np.diag(np.asarray(self.inertia_rot, float)),
                                        gravity=float(np.asarray(self.g, float).reshape(12)[11]), device=self._dev())
                self.mpc_status = int(st[0])
                f = forces[:, :, 0].cpu().numpy().T  # (12, horizon) like sol.value(controls), kalman_filter.py:152
        self.p_mpc[:] = np.asarray(p, float).reshape(12, 1)                 # kalman_filter.py:141-146
        self.body_mpc[:, 0] = np.asarray(self.x, float).reshape(12)
        self.body_mpc[:, 1:] = np.asarray(body_ref, float).reshape(12, -1)
        self.contact_mpc[:] = np.asarray(cur_contact, float).reshape(4, 1)
        f = np.asarray(f, float)
        self.f = f if f.ndim == 2 and f.shape[1] > 1 else f.reshape(12, 1)
        f0 = self.f[:, 0].reshape(12)
        body_ref = np.asarray(body_ref, float)
        br = body_ref[:, 0] if body_ref.ndim == 2 and body_ref.shape[1] > 1 else body_ref.reshape(12)
        self._predict_call(p, f0, nv.COV_MPC, body_ref=br)
        self._refresh_model_matrices(br[0:3], exp_form=True)

    @staticmethod
    def _bytes(a):
        return a.tobytes() if type(a) is np.ndarray and a.dtype == _F8 else None

    def _ahead_is_valid(self):
        a, b = self._ahead, self._bytes
        return a is not None and b(self.x) == a[0] and b(self.P) == a[1] and b(self.z) == a[2] and b(self.R) == a[3]

    def update(self):
        """kalman_filter.py:164-174."""
        self._buffers()
        if not self._ahead_is_valid():  # x, P, z or R were changed after the prediction (or there was none): update on its own
            self._put_state()
            self._put("z", self.z, 10)
            self._run(nv.PHASE_UPDATE)
        self._ahead = None
        v = self._v
        self.status = int(self._st[1])
        if self.status & nv.ST_SINGULAR:  # np.linalg.inv raises for a singular S only; an indefinite one goes through (kalman_filter.py:168)
            raise np.linalg.LinAlgError("Singular matrix")
        self.K = v["upd_K"].copy().reshape(12, 10)
        self.x = v["upd_x"].copy().reshape(12, 1)
        self.P = v["upd_P"].copy().reshape(12, 12)
        self.P_trace = float(v["upd_trace"][0])
        self.K_gain = float(v["upd_kgain"][0])

    def estimate_state_mpc(self, imu, p, dp, body_ref, contact, f=None):
        """kalman_filter.py:176-182."""
        odom = self.get_odom(p, dp, contact, imu)
        self.set_measurements(imu, odom)
        self.predict_mpc(p, body_ref, contact, f=f)
        self.update()
        return self.x

    def rotation_matrix_body_world(self, thx, thy, thz):
        """kalman_filter.py:184-193 (host helper; the filter itself evaluates R on the device)."""
        a, b, c = (float(np.asarray(v).reshape(-1)[0]) for v in (thx, thy, thz))
        sa, ca, sb, cb, sc, cc = np.sin(a), np.cos(a), np.sin(b), np.cos(b), np.sin(c), np.cos(c)
        return np.array([[cc * cb, cc * (sb * sa) - sc * ca, cc * (sb * ca) + sc * sa],
                         [sc * cb, sc * (sb * sa) + cc * ca, sc * (sb * ca) - cc * sa],
                         [-sb, cb * sa, cb * ca]])

    def skew(self, x):
        """kalman_filter.py:195-198."""
        return np.array([[0, -x[2][0], x[1][0]], [x[2][0], 0, -x[0][0]], [-x[1][0], x[0][0], 0]])
