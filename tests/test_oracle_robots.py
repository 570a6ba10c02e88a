"""oracle.robots: Pinocchio is absent (parity unpinned), so the restated rigid-body algorithms are
validated three independent ways: (1) URDF tables against the parsed reference URDFs, (2) the mass
matrix against kinetic-energy (geometric Jacobian) and Lagrangian closed forms, (3) nle against
M-dot / potential-energy finite differences, derivatives against central differences."""
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import robots

URDF = json.load(open(os.path.join(GOLDEN, 'urdf_tables.json')))
NAMES = dict(manipulator='planar_manipulator_3dof', double_integrator='double_integrator', ur5='ur5_robot')


@pytest.mark.parametrize('system', list(NAMES))
def test_chain_tables_match_urdf(system):
    u = URDF[NAMES[system]]
    ch = robots.CHAINS[system]
    moving = [j for j in u['joints'] if j['type'] != 'fixed']
    assert len(moving) == ch.n
    base = np.zeros(3)
    for j in u['joints']:
        if j['type'] == 'fixed' and j['child'] != 'EE':
            base += np.array(j['xyz'])
    for i, j in enumerate(moving):
        assert ch.jtype[i] == {'revolute': 'R', 'prismatic': 'P'}[j['type']]
        assert j['axis'][ch.axis[i]] == 1.0 and sum(abs(x) for x in j['axis']) == 1.0
        exp_p = np.array(j['xyz']) + (base if i == 0 else 0)
        np.testing.assert_array_equal(ch.p[i], exp_p)
        np.testing.assert_array_equal(ch.Rfix[i], robots._rpy(*j['rpy']))
        link = u['links'][j['child']]
        if j['child'] == 'Sy':                      # DI: the unit mass sits on the welded EE link
            link = u['links']['EE']
        if link is None:
            assert ch.mass[i] == 0
            continue
        assert ch.mass[i] == link['mass']
        np.testing.assert_array_equal(ch.com[i], link['com'])
        ixx, iyy, izz, ixy, ixz, iyz = link['inertia']
        np.testing.assert_array_equal(ch.I[i], [[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
        assert link['com_rpy'] == [0, 0, 0]
    ee = [j for j in u['joints'] if j['child'] == 'EE'][0]
    np.testing.assert_array_equal(ch.ee_xyz, ee['xyz'])


def _mass_matrix_from_kinetic_energy(ch, q, h=1e-6):
    """M = sum_i m_i Jv_i^T Jv_i + Jw_i^T (R I R^T) Jw_i with numerically differentiated FK."""
    n = ch.n
    M = np.zeros((n, n))
    frames0, _ = ch.fk(q)
    for i in range(n):
        Jv = np.zeros((3, n))
        Jw = np.zeros((3, n))
        R0, p0 = frames0[i]
        for j in range(n):
            e = np.zeros(n)
            e[j] = h
            (Rp, pp), (Rm, pm) = ch.fk(q + e)[0][i], ch.fk(q - e)[0][i]
            Jv[:, j] = ((pp + Rp @ ch.com[i]) - (pm + Rm @ ch.com[i])) / (2 * h)
            W = (Rp - Rm) / (2 * h) @ R0.T
            Jw[:, j] = [W[2, 1], W[0, 2], W[1, 0]]
        M += ch.mass[i] * Jv.T @ Jv + Jw.T @ (R0 @ ch.I[i] @ R0.T) @ Jw
    return M


def _potential(ch, q):
    frames, _ = ch.fk(q)
    return sum(ch.mass[i] * robots.GRAVITY * (frames[i][1] + frames[i][0] @ ch.com[i])[2] for i in range(ch.n))


@pytest.mark.parametrize('system', ['manipulator', 'ur5', 'double_integrator'])
def test_mass_matrix_and_nle_against_lagrangian(system):
    ch = robots.CHAINS[system]
    rng = np.random.default_rng(3)
    for _ in range(4):
        q = rng.uniform(-math.pi, math.pi, ch.n)
        v = rng.uniform(-1, 1, ch.n)
        M = ch.crba(q)
        np.testing.assert_allclose(M, M.T, atol=1e-12)
        np.testing.assert_allclose(M, _mass_matrix_from_kinetic_energy(ch, q), rtol=1e-6, atol=1e-7)
        # Lagrange: nle = Mdot v - 0.5 d(v'Mv)/dq + dU/dq
        h = 1e-6
        dM = [(ch.crba(q + h * np.eye(ch.n)[k]) - ch.crba(q - h * np.eye(ch.n)[k])) / (2 * h) for k in range(ch.n)]
        Mdot = sum(dM[k] * v[k] for k in range(ch.n))
        dT = np.array([0.5 * v @ dM[k] @ v for k in range(ch.n)])
        dU = np.array([(_potential(ch, q + h * np.eye(ch.n)[k]) - _potential(ch, q - h * np.eye(ch.n)[k])) / (2 * h) for k in range(ch.n)])
        np.testing.assert_allclose(ch.nle(q, v), Mdot @ v - dT + dU, rtol=1e-6, atol=1e-6)


def test_planar_3r_closed_form():
    """Closed form of SURVEY.md section 8a (derived there with sympy) and its known answer."""
    ch = robots.MANIPULATOR
    q = np.array([0.3, -0.7, 1.1])
    v = np.array([0.5, -0.2, 0.4])
    u = np.array([10.0, -5.0, 2.0])
    c2, c3, c23 = math.cos(q[1]), math.cos(q[2]), math.cos(q[1] + q[2])
    I = 16.666666666666668
    a = 3 * I + 0.5 * (25 + 125 + 225)                    # = 237.5
    M = np.array([[a + 150 * c2 + 50 * c3 + 50 * c23, 0, 0], [0, 0, 0], [0, 0, 0]])
    np.testing.assert_allclose(ch.crba(q)[0, 0], M[0, 0], rtol=1e-13)
    np.testing.assert_allclose(ch.nle(q, v), [-18.979195901319, -18.557290596892, 4.439081199567], rtol=1e-11)
    np.testing.assert_allclose(ch.forward_dynamics(q, v, u), [0.080045604099, 0.094780798639, -0.389618439764], rtol=1e-9)
    np.testing.assert_allclose(ch.ee_position(q), [19.41239670413, 5.503195515904, 0], rtol=1e-11)


@pytest.mark.parametrize('system', ['manipulator', 'ur5'])
def test_aba_derivatives_against_central_differences(system):
    ch = robots.CHAINS[system]
    rng = np.random.default_rng(5)
    q = rng.uniform(-2, 2, ch.n)
    v = rng.uniform(-1, 1, ch.n)
    tau = rng.uniform(-10, 10, ch.n)
    dq, dv, Minv = ch.aba_derivatives(q, v, tau)
    h = 1e-6
    E = np.eye(ch.n)
    fdq = np.array([(ch.forward_dynamics(q + h * E[j], v, tau) - ch.forward_dynamics(q - h * E[j], v, tau)) / (2 * h) for j in range(ch.n)]).T
    fdv = np.array([(ch.forward_dynamics(q, v + h * E[j], tau) - ch.forward_dynamics(q, v - h * E[j], tau)) / (2 * h) for j in range(ch.n)]).T
    np.testing.assert_allclose(dq, fdq, rtol=1e-5, atol=1e-5 * np.abs(fdq).max())
    np.testing.assert_allclose(dv, fdv, rtol=1e-5, atol=1e-5 * np.abs(fdv).max())
    np.testing.assert_allclose(Minv @ ch.crba(q), np.eye(ch.n), atol=1e-9)


def test_double_integrator_is_unit_mass_point():
    ch = robots.DOUBLE_INTEGRATOR
    np.testing.assert_allclose(ch.crba(np.array([1.3, -2.0])), np.eye(2), atol=1e-15)
    np.testing.assert_allclose(ch.nle(np.array([1.3, -2.0]), np.array([0.4, 5.0])), 0, atol=1e-15)
    np.testing.assert_allclose(ch.ee_position(np.array([1.3, -2.0])), [1.3, -2.0, 0.0])
