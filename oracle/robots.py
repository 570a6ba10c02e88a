"""Rigid-body algorithms for serial chains (NumPy fp64) -- ORACLE, test infrastructure only.

Restates what the reference obtains from Pinocchio (pin3x-jnrh2023==2.9.2, NOT present in
/root/reference -> "parity unpinned", see oracle/__init__.py):

* ``computeAllTerms``  -> M(q) (CRBA), nle(q, v)            robot_utils.py:353-356
* ``integrate``        -> q + v*dt for R^n joints            robot_utils.py:402
* ``computeABADerivatives`` -> ddq_dq, ddq_dv, Minv          environment.py:100,120-126
* ``framePlacement(q, 'EE')``                                environment.py:150-154

Conventions follow SURVEY.md A.7: joint placement R = Rz(yaw) Ry(pitch) Rx(roll),
axis-aligned revolute/prismatic joints, link inertia (mass, com, I about com), fixed
children merged into the parent, gravity (0, 0, -9.81).

The algorithm is a classical 3-vector recursive Newton-Euler written so that it also
accepts complex inputs: exact derivatives are taken by complex-step differentiation.

Chain tables below were extracted from /root/reference/urdf/*.urdf by
tests/golden/make_golden.py::urdf_tables (planar_manipulator_3dof.urdf:24-98,
double_integrator.urdf:6-64, ur5_robot.urdf:28-212).
"""
import math
import numpy as np

GRAVITY = 9.81


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


class Chain:
    """Serial kinematic chain. joints: list of dict(type 'R'|'P', axis 0|1|2, xyz, rpy,
    mass, com, inertia(ixx, iyy, izz, ixy, ixz, iyz)); ee: dict(xyz, rpy) on the last link;
    base_xyz: fixed offset world -> first joint parent."""

    def __init__(self, name, joints, ee_xyz, ee_rpy=(0, 0, 0), base_xyz=(0, 0, 0)):
        self.name = name
        self.n = len(joints)
        self.jtype = [j['type'] for j in joints]
        self.axis = [j['axis'] for j in joints]
        self.p = [np.array(j['xyz'], dtype=float) for j in joints]
        self.p[0] = self.p[0] + np.array(base_xyz, dtype=float)
        self.Rfix = [_rpy(*j['rpy']) for j in joints]
        self.mass = [float(j['mass']) for j in joints]
        self.com = [np.array(j['com'], dtype=float) for j in joints]
        self.I = []
        for j in joints:
            ixx, iyy, izz, ixy, ixz, iyz = j['inertia']
            self.I.append(np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]], dtype=float))
        self.ee_xyz = np.array(ee_xyz, dtype=float)
        self.ee_R = _rpy(*ee_rpy)

    # -- kinematics -------------------------------------------------------------------
    def _joint_rot(self, i, qi):
        c, s = np.cos(qi), np.sin(qi)
        one, zero = c * 0 + 1, c * 0
        a = self.axis[i]
        if a == 0:
            return np.array([[one, zero, zero], [zero, c, -s], [zero, s, c]])
        if a == 1:
            return np.array([[c, zero, s], [zero, one, zero], [-s, zero, c]])
        return np.array([[c, -s, zero], [s, c, zero], [zero, zero, one]])

    def _unit(self, i, dtype):
        e = np.zeros(3, dtype=dtype)
        e[self.axis[i]] = 1
        return e

    def fk(self, q):
        """World placement (R, p) of every joint frame and of the EE frame."""
        q = np.asarray(q)
        R = np.eye(3, dtype=q.dtype)
        p = np.zeros(3, dtype=q.dtype)
        frames = []
        for i in range(self.n):
            if self.jtype[i] == 'R':
                p = p + R @ self.p[i]
                R = R @ self.Rfix[i] @ self._joint_rot(i, q[i])
            else:
                Rn = R @ self.Rfix[i]
                p = p + R @ self.p[i] + Rn @ (self._unit(i, q.dtype) * q[i])
                R = Rn
            frames.append((R, p))
        return frames, (R @ self.ee_R, p + R @ self.ee_xyz)

    def ee_position(self, q):
        return self.fk(q)[1][1]

    # -- dynamics ---------------------------------------------------------------------
    def rnea(self, q, v, a, gravity=GRAVITY):
        """tau = M(q) a + nle(q, v)  (inverse dynamics)."""
        q, v, a = np.asarray(q), np.asarray(v), np.asarray(a)
        dt = np.result_type(q.dtype, v.dtype, a.dtype, float)
        n = self.n
        w = np.zeros(3, dtype=dt)
        wd = np.zeros(3, dtype=dt)
        acc = np.array([0, 0, gravity], dtype=dt)        # base accelerates upwards <=> gravity
        Rs, ps, F, N = [], [], [], []
        for i in range(n):
            e = self._unit(i, dt)
            if self.jtype[i] == 'R':
                R = self.Rfix[i] @ self._joint_rot(i, q[i])          # parent_R_child
                p = self.p[i].astype(dt)
                Rt = R.T
                wp = Rt @ w
                acc = Rt @ (acc + np.cross(wd, p) + np.cross(w, np.cross(w, p)))
                wd = Rt @ wd + np.cross(wp, e * v[i]) + e * a[i]
                w = wp + e * v[i]
            else:
                R = self.Rfix[i].astype(dt)
                p = self.p[i] + R @ (e * q[i])
                Rt = R.T
                acc0 = Rt @ (acc + np.cross(wd, p) + np.cross(w, np.cross(w, p)))
                w = Rt @ w
                wd = Rt @ wd
                acc = acc0 + 2 * np.cross(w, e * v[i]) + e * a[i]
            c = self.com[i]
            ac = acc + np.cross(wd, c) + np.cross(w, np.cross(w, c))
            F.append(self.mass[i] * ac)
            N.append(self.I[i] @ wd + np.cross(w, self.I[i] @ w))
            Rs.append(R)
            ps.append(p)
        tau = np.zeros(n, dtype=dt)
        f = np.zeros(3, dtype=dt)
        nn = np.zeros(3, dtype=dt)
        for i in range(n - 1, -1, -1):
            # wrench of the child (expressed in child frame i+1) moved into frame i
            if i < n - 1:
                fc = Rs[i + 1] @ f
                nn = Rs[i + 1] @ nn + np.cross(ps[i + 1], fc)
                f = fc
            f = f + F[i]
            nn = nn + N[i] + np.cross(self.com[i], F[i])
            e = self._unit(i, dt)
            tau[i] = (e @ nn) if self.jtype[i] == 'R' else (e @ f)
        return tau

    def nle(self, q, v):
        return self.rnea(q, v, np.zeros(self.n))

    def crba(self, q):
        q = np.asarray(q)
        n = self.n
        M = np.zeros((n, n), dtype=np.result_type(q.dtype, float))
        z = np.zeros(n)
        for j in range(n):
            e = np.zeros(n)
            e[j] = 1.0
            M[:, j] = self.rnea(q, z, e, gravity=0.0)
        return M

    def forward_dynamics(self, q, v, tau):
        """ddq = M^-1 (tau - nle)  -- RobotSimulator.step Euler branch, robot_utils.py:399-401
        (tau_c = 0, no contacts)."""
        return np.linalg.solve(self.crba(q), np.asarray(tau) - self.nle(q, v))

    def minv(self, q):
        """Minv as aba_derivatives returns it (symmetrised inverse of the CRBA matrix) without the derivative passes:
        what Env.derivative needs (environment.py:100-104 reads only data.Minv)."""
        Minv = np.linalg.inv(self.crba(np.asarray(q, dtype=float)))
        return 0.5 * (Minv + Minv.T)

    def aba_derivatives(self, q, v, tau, h=1e-30):
        """(ddq_dq, ddq_dv, Minv) as pin.computeABADerivatives leaves them in data
        (environment.py:120-126).  Exact via complex step."""
        q = np.asarray(q, dtype=float)
        v = np.asarray(v, dtype=float)
        tau = np.asarray(tau, dtype=float)
        n = self.n
        ddq_dq = np.zeros((n, n))
        ddq_dv = np.zeros((n, n))
        for j in range(n):
            qc = q.astype(complex)
            qc[j] += 1j * h
            ddq_dq[:, j] = np.imag(self.forward_dynamics(qc, v, tau)) / h
            vc = v.astype(complex)
            vc[j] += 1j * h
            ddq_dv[:, j] = np.imag(self.forward_dynamics(q, vc, tau)) / h
        return ddq_dq, ddq_dv, self.minv(q)


_Z3 = (0.0, 0.0, 0.0)
_HP = 1.57079632679     # the URDF's own truncated pi/2 (ur5_robot.urdf:62,122,210)

MANIPULATOR = Chain(
    'manipulator',
    [dict(type='R', axis=2, xyz=_Z3, rpy=_Z3, mass=0.5, com=(5, 0, 0),
          inertia=(16.666666666666668, 0.0, 16.666666666666668, 0, 0, 0)),
     dict(type='R', axis=2, xyz=(10, 0, 0), rpy=_Z3, mass=0.5, com=(5, 0, 0),
          inertia=(16.666666666666668, 0.0, 16.666666666666668, 0, 0, 0)),
     dict(type='R', axis=2, xyz=(10, 0, 0), rpy=_Z3, mass=0.5, com=(5, 0, 0),
          inertia=(16.666666666666668, 0.0, 16.666666666666668, 0, 0, 0))],
    ee_xyz=(10, 0, 0), base_xyz=(-7, 0, 0))

DOUBLE_INTEGRATOR = Chain(
    'double_integrator',
    [dict(type='P', axis=0, xyz=_Z3, rpy=_Z3, mass=0.0, com=_Z3, inertia=(0, 0, 0, 0, 0, 0)),
     dict(type='P', axis=1, xyz=_Z3, rpy=_Z3, mass=1.0, com=_Z3, inertia=(0, 0, 1, 0, 0, 0))],
    ee_xyz=_Z3)

UR5 = Chain(
    'ur5',
    [dict(type='R', axis=2, xyz=(0, 0, 0.089159), rpy=_Z3, mass=3.7, com=_Z3,
          inertia=(0.010267495893, 0.010267495893, 0.00666, 0, 0, 0)),
     dict(type='R', axis=1, xyz=(0, 0.13585, 0), rpy=(0, _HP, 0), mass=8.393, com=(0, 0, 0.28),
          inertia=(0.22689067591, 0.22689067591, 0.0151074, 0, 0, 0)),
     dict(type='R', axis=1, xyz=(0, -0.1197, 0.425), rpy=_Z3, mass=2.275, com=(0, 0, 0.25),
          inertia=(0.049443313556, 0.049443313556, 0.004095, 0, 0, 0)),
     dict(type='R', axis=1, xyz=(0, 0, 0.39225), rpy=(0, _HP, 0), mass=1.219, com=_Z3,
          inertia=(0.111172755531, 0.111172755531, 0.21942, 0, 0, 0)),
     dict(type='R', axis=2, xyz=(0, 0.093, 0), rpy=_Z3, mass=1.219, com=_Z3,
          inertia=(0.111172755531, 0.111172755531, 0.21942, 0, 0, 0)),
     dict(type='R', axis=1, xyz=(0, 0, 0.09465), rpy=_Z3, mass=0.1879, com=_Z3,
          inertia=(0.0171364731454, 0.0171364731454, 0.033822, 0, 0, 0))],
    ee_xyz=(0, 0.0823, 0), ee_rpy=(0, 0, _HP))

CHAINS = {'manipulator': MANIPULATOR, 'double_integrator': DOUBLE_INTEGRATOR, 'ur5': UR5}
