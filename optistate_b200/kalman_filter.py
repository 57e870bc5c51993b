"""Drop-in for the reference's `Kalman_Filter` (/root/reference/kalman_filter/kalman_filter.py:7-202).

Same constructor (no arguments, reads settings.INITIAL_PARAMS), same attributes (NumPy arrays the caller may
read and overwrite between calls: x, z, H, P, Q, R, P_trace, F, B, g, dt, m, inertia_rot, x_model, K, K_gain,
F_d, B_d), same methods and the same side effects:
  * `predict` rotates the caller's `p` into the world frame in place (force_controller.py:274-277);
  * `set_measurements` writes into `self.z` in place (kalman_filter.py:113-117);
  * `predict` / `update` rebind `self.x` / `self.P` to new arrays (kalman_filter.py:133-135,169-172), so the
    STARTING_STATE / Q aliases are only mutated by callers that assign through `KF.x[:] = ...`.
Every numerical method runs on the GPU through the C ABI (one trajectory, one step, the JOINT kernel which keeps
the reference's operand order on a full non-symmetrised P); there is no CPU implementation behind it.
For many trajectories or many steps use `optistate_b200.kf_batch`, which runs the same arithmetic in one launch.

Errors follow the reference: `numpy.linalg.LinAlgError` when S is not invertible (kalman_filter.py:168) and
`ValueError` when no foot is in stance (kalman_filter.py:97-103).

The forces inside `predict_mpc` come from the reference's convex MPC (CasADi + qpOASES, force_controller.py:15-225).  Here
they are solved for on the device by optistate_b200.mpc.mpc_forces (same QP, interior point + active-set polish; parity
with qpOASES unpinned, see there), unless forces are passed explicitly (`f=`) or an injectable `force_provider` is set.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nv
from .settings import INITIAL_PARAMS

_SEL = [0, 1, 2, 5, 6, 7, 8, 9, 10, 11]


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

    def _consts(self):
        I = np.asarray(self.inertia_rot, float)
        return dict(dt=float(self.dt), mass=float(self.m), inertia0=float(I[0, 0]), inertia1=float(I[1, 1]),
                    inertia2=float(I[2, 2]), gravity=float(np.asarray(self.g, float).reshape(12)[11]))

    @staticmethod
    def _col(a, n):
        return np.asarray(a, dtype=np.float64).reshape(n)

    # One pinned host buffer and one device buffer each way, allocated once per instance: a call is then one asynchronous
    # host->device copy, one launch, two asynchronous device->host copies (values, status) and ONE stream synchronisation,
    # issued by the binding (kf_host_call) below the Python dispatcher - no allocation, no pageable staging copy, no
    # device-side concatenation (which cost the first version ~110 us per call).
    _IN_MAX = 12 + 144 + 144 + 100 + 12 + 12 + 12 + 10 + 6 + 12 + 4
    _OUT_SIZES = {"x_final": 12, "P_final": 144, "K_final": 120, "p_world_steps": 12, "p_trace_steps": 1, "k_gain_steps": 1, "odom": 4}
    _OUT_MAX = 12 + 144 + 120 + 12 + 1 + 1 + 4

    def _buffers(self):
        dev = self._dev()
        if getattr(self, "_buf_dev", None) != dev:
            self._buf_dev = dev
            self._h_in = torch.empty(self._IN_MAX, dtype=torch.float64).pin_memory()
            self._h_in_np = self._h_in.numpy()
            self._d_in = torch.empty(self._IN_MAX, dtype=torch.float64, device=dev)
            self._d_out = torch.empty(self._OUT_MAX, dtype=torch.float64, device=dev)
            self._h_out = torch.empty(self._OUT_MAX, dtype=torch.float64).pin_memory()
            self._h_out_np = self._h_out.numpy()
            self._d_status = torch.zeros(1, dtype=torch.int32, device=dev)
            self._h_status = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._h_status_np = self._h_status.numpy()
        return dev

    def _launch(self, host_in, want, cfg, consts):
        """Packs `host_in` (name -> array, in order) into the pinned buffer; the binding uploads, launches (kf_batch with `cfg`, or
        kf_measure when cfg is empty), downloads the `want` outputs and the status word and synchronises.  Returns {name: NumPy copy}."""
        self._buffers()
        in_layout, off = [], 0
        for k, v in host_in.items():
            n = v.size
            self._h_in_np[off:off + n] = v.reshape(-1)
            in_layout.append((k, n))
            off += n
        out_layout = [(k, self._OUT_SIZES[k]) for k in want]
        nv.check(int(nv.ext().kf_host_call(cfg, consts, in_layout, out_layout, self._h_in, self._d_in, self._d_out, self._h_out,
                                            self._d_status, self._h_status)), "optistate_kf_batch" if cfg else "optistate_kf_measure")
        self.status = int(self._h_status_np[0])
        res, off = {}, 0
        for k, n in out_layout:
            res[k] = self._h_out_np[off:off + n].copy()
            off += n
        return res

    def _step(self, phases, cov_model, host_in, want):
        """One trajectory, one step through optistate_kf_batch (JOINT)."""
        cfg = dict(dtype=nv.F64, algo=nv.ALGO_JOINT, cov_model=cov_model, phases=phases, n_traj=1, n_steps=1, n_streams=1,
                   x0_per_traj=0, p0_kind=nv.MAT_DENSE, q_kind=nv.MAT_DENSE, r_kind=nv.MAT_DENSE)
        return self._launch(host_in, want, cfg, self._consts())

    def _state_in(self):
        return {"x0": self._col(self.x, 12), "P0": np.asarray(self.P, dtype=np.float64).reshape(144),
                "Q": np.asarray(self.Q, dtype=np.float64).reshape(144), "R": np.asarray(self.R, dtype=np.float64).reshape(100)}

    # ------------------------------------------------------------------ reference surface
    def get_odom(self, p_cur, dp_cur, contact_cur, imu):
        """kalman_filter.py:79-105 -> (4,1) array [z, vx, vy, vz]."""
        host_in = {"imu": self._col(imu, 6), "p": self._col(p_cur, 12), "dp": self._col(dp_cur, 12), "contact": self._col(contact_cur, 4)}
        r = self._launch(host_in, ["odom"], {}, {})
        if self.status & nv.ST_ALL_SWING:
            raise ValueError("setting an array element with a sequence. The requested array has an inhomogeneous shape "
                             "(all four feet in swing; the reference fails here too, kalman_filter.py:97-103)")
        return r["odom"].reshape(4, 1)

    def set_measurements(self, imu, odom):
        """kalman_filter.py:108-117 (in-place scatter into self.z)."""
        imu, odom = np.asarray(imu), np.asarray(odom)
        self.z[0:3] = imu[0:3].reshape(3, 1)
        self.z[4:7] = imu[3:6].reshape(3, 1)
        self.z[3] = odom[0]
        self.z[7:10] = odom[1:].reshape(3, 1)

    def _refresh_model_matrices(self, angles, exp_form):
        R = self.rotation_matrix_body_world(angles[0], angles[1], angles[2])
        self.F[0:3, 6:9] = np.transpose(R)
        self.F_d = np.exp(self.dt * self.F) if exp_form else self.identity_large + self.dt * self.F
        self.B_d = self.dt * self.B

    def predict(self, p, f):
        """kalman_filter.py:119-138; p (12,1) is rotated into the world frame in place."""
        prior = self._col(self.x, 12).copy()
        host_in = dict(self._state_in(), p=self._col(p, 12), f=self._col(f, 12))
        r = self._step(nv.PHASE_PREDICT, nv.COV_PREDICT, host_in, ["x_final", "P_final", "p_world_steps", "p_trace_steps"])
        self._refresh_model_matrices(prior[0:3], exp_form=False)
        p[...] = r["p_world_steps"].reshape(np.shape(p))
        self.x = r["x_final"].reshape(12, 1)
        self.P = r["P_final"].reshape(12, 12)
        self.x_model = self.x.copy()
        self.P_trace = float(r["p_trace_steps"][0])

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
        host_in = dict(self._state_in(), p=self._col(p, 12), f=f0, body_ref=br.reshape(12))
        r = self._step(nv.PHASE_PREDICT, nv.COV_MPC, host_in, ["x_final", "P_final", "p_world_steps", "p_trace_steps"])
        self._refresh_model_matrices(br[0:3], exp_form=True)
        p[...] = r["p_world_steps"].reshape(np.shape(p))
        self.x = r["x_final"].reshape(12, 1)
        self.P = r["P_final"].reshape(12, 12)
        self.x_model = self.x.copy()

    def update(self):
        """kalman_filter.py:164-174."""
        host_in = dict(self._state_in(), z_in=self._col(self.z, 10))
        r = self._step(nv.PHASE_UPDATE, nv.COV_PREDICT, host_in, ["x_final", "P_final", "K_final", "p_trace_steps", "k_gain_steps"])
        if self.status & nv.ST_SINGULAR:  # np.linalg.inv raises for a singular S only; an indefinite one goes through (kalman_filter.py:168)
            raise np.linalg.LinAlgError("Singular matrix")
        self.K = r["K_final"].reshape(12, 10)
        self.x = r["x_final"].reshape(12, 1)
        self.P = r["P_final"].reshape(12, 12)
        self.P_trace = float(r["p_trace_steps"][0])
        self.K_gain = float(r["k_gain_steps"][0])

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
