"""Rigid-body parameters of the Pinocchio-backed systems, as serial-chain tables.

The reference builds these with ``RobotWrapper.BuildFromURDF`` at conf import time
(conf_manipulator.py:157-160, conf_double_integrator.py, conf_ur5.py).  Here they are plain data,
packed into the ``cacto_chain`` POD that the kernels read (include/cacto_b200.h).  Conventions
(SURVEY.md A.7): joint placement = URDF <origin xyz rpy> with R = Rz(yaw) Ry(pitch) Rx(roll); fixed
joints are folded into their parent (``base`` offset, ``ee`` frame); link inertia = (mass, COM,
[ixx iyy izz ixy ixz iyz] about the COM).

tests/test_robots_tables.py checks every number against the parsed URDFs
(tests/golden/urdf_tables.json).
"""
import math

REVOLUTE, PRISMATIC = 0, 1
X, Y, Z = 0, 1, 2
_HALF_PI_URDF = 1.57079632679        # the URDF's own literal (ur5_robot.urdf:62,122,210)


def _joint(kind, axis, xyz=(0.0, 0.0, 0.0), rpy=(0.0, 0.0, 0.0), mass=0.0, com=(0.0, 0.0, 0.0), inertia=(0.0,) * 6):
    return dict(kind=kind, axis=axis, xyz=tuple(map(float, xyz)), rpy=tuple(map(float, rpy)), mass=float(mass),
                com=tuple(map(float, com)), inertia=tuple(map(float, inertia)))


_I_LINK = (16.666666666666668, 0.0, 16.666666666666668, 0.0, 0.0, 0.0)

CHAINS = {
    # urdf/planar_manipulator_3dof.urdf: three revolute-z joints 10 apart, base welded at (-7, 0, 0)
    'manipulator': dict(
        base=(-7.0, 0.0, 0.0),
        joints=[_joint(REVOLUTE, Z, mass=0.5, com=(5, 0, 0), inertia=_I_LINK),
                _joint(REVOLUTE, Z, xyz=(10, 0, 0), mass=0.5, com=(5, 0, 0), inertia=_I_LINK),
                _joint(REVOLUTE, Z, xyz=(10, 0, 0), mass=0.5, com=(5, 0, 0), inertia=_I_LINK)],
        ee=(10.0, 0.0, 0.0)),
    # urdf/double_integrator.urdf: prismatic x then prismatic y carrying a unit point mass
    'double_integrator': dict(
        base=(0.0, 0.0, 0.0),
        joints=[_joint(PRISMATIC, X), _joint(PRISMATIC, Y, mass=1.0, inertia=(0, 0, 1, 0, 0, 0))],
        ee=(0.0, 0.0, 0.0)),
    # urdf/ur5_robot.urdf
    'ur5': dict(
        base=(0.0, 0.0, 0.0),
        joints=[_joint(REVOLUTE, Z, xyz=(0, 0, 0.089159), mass=3.7, inertia=(0.010267495893, 0.010267495893, 0.00666, 0, 0, 0)),
                _joint(REVOLUTE, Y, xyz=(0, 0.13585, 0), rpy=(0, _HALF_PI_URDF, 0), mass=8.393, com=(0, 0, 0.28),
                       inertia=(0.22689067591, 0.22689067591, 0.0151074, 0, 0, 0)),
                _joint(REVOLUTE, Y, xyz=(0, -0.1197, 0.425), mass=2.275, com=(0, 0, 0.25),
                       inertia=(0.049443313556, 0.049443313556, 0.004095, 0, 0, 0)),
                _joint(REVOLUTE, Y, xyz=(0, 0, 0.39225), rpy=(0, _HALF_PI_URDF, 0), mass=1.219,
                       inertia=(0.111172755531, 0.111172755531, 0.21942, 0, 0, 0)),
                _joint(REVOLUTE, Z, xyz=(0, 0.093, 0), mass=1.219, inertia=(0.111172755531, 0.111172755531, 0.21942, 0, 0, 0)),
                _joint(REVOLUTE, Y, xyz=(0, 0, 0.09465), mass=0.1879, inertia=(0.0171364731454, 0.0171364731454, 0.033822, 0, 0, 0))],
        ee=(0.0, 0.0823, 0.0)),
}

GRAVITY = 9.81


def rpy_matrix(roll, pitch, yaw):
    """Row-major 3x3 of Rz(yaw) Ry(pitch) Rx(roll)."""
    cr, sr, cp, sp, cy, sy = math.cos(roll), math.sin(roll), math.cos(pitch), math.sin(pitch), math.cos(yaw), math.sin(yaw)
    return [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
            sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
            -sp, cp * sr, cp * cr]
