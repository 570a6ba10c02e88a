"""CPU restatement of ``TO_Casadi.backward_pass`` (TO.py:119-202) -- ORACLE, test infrastructure only.

The reference differentiates ``running_cost = -runningSingleModel.cost(x, u)`` (TO.py:147-164) with CasADi
(``casadi==3.6.3``, absent here).  Pinning: for the single integrator, the car and car_park (the systems whose models need no
Pinocchio) the reference's OWN ``TO_Casadi.backward_pass`` + ``*_CAMS`` cost models + ``Env.augmented_derivative`` were
executed unmodified in the build container on a small symbolic stand-in for CasADi (tests/golden/_casadi_stub.py: expression
graph, exact hyper-dual derivatives) -> tests/golden/bp_cases.npz; this module reproduces those V_x to 1e-9 and the cost
values to 1e-12 (tests/test_oracle_backward.py).  Double integrator, manipulator, UR5: PARITY UNPINNED (their CAMS
models need pinocchio.casadi); covered by the checks listed below.  ``-cost`` is the reward of environment.py with the
running / terminal weights (environment_TO.py:90-111 SI, :208-234 DI, :339-360 car, :479-503 car_park, :605-631
manipulator, :735-765 UR5 -- each the negative of the matching ``Env.reward``, with the bounded control cost
``a^2 + w_b (a/u_max)^10`` for every system including UR5, quirk Q8).  Here the same expression is evaluated on
hyper-dual numbers (value, two first-order parts, one mixed second-order part), which gives exact gradients and
Hessians without a symbolic engine; the recursion then follows TO.py:166-200 line by line with NumPy
(``np.linalg.pinv`` included).  Checks in tests/test_oracle_backward.py: the generic reward equals the pinned
``oracle.systems`` reward on floats, derivatives equal central finite differences, and for the double integrator
(linear dynamics) the result equals the finite-difference gradient of the closed-loop-free value recursion.
"""
import math

import numpy as np

from . import systems as osys


class HD:
    """Hyper-dual number v + a e1 + b e2 + ab e1 e2 (e1^2 = e2^2 = 0): f(x + e1 + e2) carries f, f', f', f''."""
    __slots__ = ('v', 'a', 'b', 'ab')

    def __init__(self, v, a=0.0, b=0.0, ab=0.0):
        self.v, self.a, self.b, self.ab = float(v), float(a), float(b), float(ab)

    @staticmethod
    def lift(x):
        return x if isinstance(x, HD) else HD(x)

    def _un(self, f0, f1, f2):
        return HD(f0, f1 * self.a, f1 * self.b, f1 * self.ab + f2 * self.a * self.b)

    def __add__(self, o):
        o = HD.lift(o)
        return HD(self.v + o.v, self.a + o.a, self.b + o.b, self.ab + o.ab)
    __radd__ = __add__

    def __neg__(self):
        return HD(-self.v, -self.a, -self.b, -self.ab)

    def __sub__(self, o):
        return self + (-HD.lift(o))

    def __rsub__(self, o):
        return HD.lift(o) + (-self)

    def __mul__(self, o):
        o = HD.lift(o)
        return HD(self.v * o.v, self.v * o.a + self.a * o.v, self.v * o.b + self.b * o.v,
                  self.v * o.ab + self.a * o.b + self.b * o.a + self.ab * o.v)
    __rmul__ = __mul__

    def recip(self):
        return self._un(1.0 / self.v, -1.0 / self.v ** 2, 2.0 / self.v ** 3)

    def __truediv__(self, o):
        return self * HD.lift(o).recip()

    def __rtruediv__(self, o):
        return HD.lift(o) * self.recip()

    def __pow__(self, n):
        if isinstance(n, int) and n >= 0:
            r = HD(1.0)
            for _ in range(n):
                r = r * self
            return r
        n = float(n)
        return self._un(self.v ** n, n * self.v ** (n - 1), n * (n - 1) * self.v ** (n - 2))

    def sqrt(self):
        s = math.sqrt(self.v)
        return self._un(s, 0.5 / s, -0.25 / (s * self.v))

    def exp(self):
        e = math.exp(self.v)
        return self._un(e, e, e)

    def log(self):
        return self._un(math.log(self.v), 1.0 / self.v, -1.0 / self.v ** 2)

    def sin(self):
        return self._un(math.sin(self.v), math.cos(self.v), -math.sin(self.v))

    def cos(self):
        return self._un(math.cos(self.v), -math.sin(self.v), -math.cos(self.v))


def _f(name, x):
    return getattr(x, name)() if isinstance(x, HD) else getattr(math, name)(x)


def _ee(env, q):
    """End-effector position for float or HD joint values (environment_TO.py p_ee; oracle.robots FK for the chains)."""
    sid = env.conf.system_id
    if sid in ('single_integrator', 'car'):
        return [q[0], q[1], 0.0]
    if sid == 'car_park':                                              # environment_TO.py:449-454
        L = env.conf.L_delta / 2
        return [q[0] + _f('cos', q[2]) * L, q[1] + _f('sin', q[2]) * L, 0.0]
    qa = np.empty(env.nq, dtype=object)
    for i in range(env.nq):
        qa[i] = q[i]
    p = env.chain.ee_position(qa if any(isinstance(v, HD) for v in q[:env.nq]) else np.asarray(q[:env.nq], dtype=float))
    return [p[0], p[1], p[2]]


def _softplus_over(alpha, z):
    return _f('log', _f('exp', alpha * z) + 1) / alpha


def reward_generic(env, weights, x, u=None):
    """``-CAMS.cost_fun(x, u)`` = ``Env.reward(weights, x, u)`` with the bounded control cost; x, u floats or HD."""
    c, sid = env.conf, env.conf.system_id
    o, T = c.obs_param, env.TARGET_STATE
    p = _ee(env, x)
    dims = 3 if sid == 'ur5' else 2
    s = 0.0
    for i in range(dims):
        s = s + _f('sqrt', (p[i] - T[i]) ** 2 + 0.1) - math.sqrt(0.1) - 0.1
    peak = _softplus_over(env.alpha2, -s)
    dist = 0.0
    for i in range(dims):
        dist = dist + (p[i] - T[i]) ** 2
    u_cost = 0.0
    if u is not None:
        for i in range(c.nb_action):
            u_cost = u_cost + u[i] * u[i] + c.w_b * (u[i] / c.u_max[i]) ** 10
    if sid == 'car_park':                                              # environment_TO.py:479-503
        k = c.k_db
        ct, st = _f('cos', x[2]), _f('sin', x[2])
        obs = 0.0

        def sig(z):                                                    # 0.5 (1 + z / sqrt(1 + z^2)), SURVEY A.1
            return 0.5 * (1 + z / _f('sqrt', 1 + z * z))
        for kk in range(3):
            xc, yc, Wx, Wy = o[2 * kk], o[2 * kk + 1], o[6 + 2 * kk], o[7 + 2 * kk]
            for bx, by in np.asarray(c.check_points_BF):
                px, py = ct * bx - st * by + p[0], st * bx + ct * by + p[1]
                obs = obs + (sig((py - yc + Wy / 2) * k) * (1 - sig((py - yc - Wy / 2) * k))
                             * sig((px - xc + Wx / 2) * k) * (1 - sig((px - xc - Wx / 2) * k)))
        r = -weights[0] * dist + weights[1] * peak - weights[2] * x[3] ** 2 - weights[3] * obs - weights[6] * u_cost + env.offset
        return env.scale * r
    ell = []
    for kk in range(3):
        if sid == 'ur5':
            e = ((p[0] - o[3 * kk]) ** 2 / (o[9 + 3 * kk] / 2) ** 2 + (p[1] - o[3 * kk + 1]) ** 2 / (o[10 + 3 * kk] / 2) ** 2
                 + (p[2] - o[3 * kk + 2]) ** 2 / (o[11 + 3 * kk] / 2) ** 2 - 1.0)
        else:
            e = (p[0] - o[2 * kk]) ** 2 / (o[6 + 2 * kk] / 2) ** 2 + (p[1] - o[2 * kk + 1]) ** 2 / (o[7 + 2 * kk] / 2) ** 2 - 1.0
        ell.append(_softplus_over(env.alpha, -e))
    vel = 0.0
    if sid in ('manipulator', 'ur5'):
        for i in range(env.nq, env.nx):
            vel = vel + x[i] ** 2
    r = (-weights[0] * dist + weights[1] * peak - weights[2] * vel - weights[3] * ell[0] - weights[4] * ell[1]
         - weights[5] * ell[2] - weights[6] * u_cost + env.offset)
    return env.scale * r


def reward_x_derivatives(env, weights, x):
    """(l_x[n], l_xx[n, n]) of x -> reward_generic(env, weights, x) by hyper-dual evaluation (TO.py:150-153)."""
    n = env.nx
    g, H = np.zeros(n), np.zeros((n, n))
    for i in range(n):
        for j in range(i, n):
            xs = [HD(x[k], 1.0 if k == i else 0.0, 1.0 if k == j else 0.0) for k in range(n)]
            r = reward_generic(env, weights, xs)
            r = HD.lift(r)
            H[i, j] = H[j, i] = r.ab
            if j == i:
                g[i] = r.a
    return g, H


def reward_u_derivatives(env, weights, u):
    """(l_u[m], l_uu[m, m]) of the control part: -scale w6 sum(a^2 + w_b (a/u_max)^10) (TO.py:151, environment_TO.py:84-88)."""
    c = env.conf
    u = np.asarray(u, dtype=float)
    um = np.asarray(c.u_max, dtype=float)
    g = -env.scale * weights[6] * (2 * u + 10 * c.w_b * u ** 9 / um ** 10)
    H = np.diag(-env.scale * weights[6] * (2 + 90 * c.w_b * u ** 8 / um ** 10))
    return g, H


def backward_pass(env, T, TO_states, TO_controls, mu=1e-9):
    """TO.py:119-202: V_x[T, n+1] (last column = 0) of the reward-to-go along (TO_states[T, n], TO_controls[T-1, m])."""
    c = env.conf
    n, m = c.nb_state - 1, c.nb_action
    X = np.asarray(TO_states, dtype=float)[:T, :n]
    U = np.asarray(TO_controls, dtype=float).reshape(-1, m)[:T - 1]
    V_xx = np.zeros((T, n, n))
    V_x = np.zeros((T, n + 1))
    l_x, l_xx = reward_x_derivatives(env, c.cost_weights_terminal, X[-1])          # :172-174
    V_xx[T - 1], V_x[T - 1, :-1] = l_xx, l_x
    for i in range(T - 2, -1, -1):                                                  # :176-200
        A, B = env.augmented_derivative(np.append(X[i], 0.0), U[i])
        l_x, l_xx = reward_x_derivatives(env, c.cost_weights_running, X[i])
        l_u, l_uu = reward_u_derivatives(env, c.cost_weights_running, U[i])
        l_xu = np.zeros((n, m))                                                     # the cost is separable in x and u
        Q_x = l_x + A.T @ V_x[i + 1, :-1]
        Q_u = l_u + B.T @ V_x[i + 1, :-1]
        Q_xx = l_xx + A.T @ V_xx[i + 1] @ A
        Q_uu = l_uu + B.T @ V_xx[i + 1] @ B
        Q_xu = l_xu + A.T @ V_xx[i + 1] @ B
        Qbar_uu_pinv = np.linalg.pinv(Q_uu + mu * np.identity(m))
        V_x[i, :-1] = Q_x - Q_xu @ Qbar_uu_pinv @ Q_u
        V_xx[i] = Q_xx - Q_xu @ Qbar_uu_pinv @ Q_xu.T
    return V_x
