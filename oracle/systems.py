"""CPU restatement of the reference's system models (environment.py) -- ORACLE, test only.

One class per system with the reference's method names.  fp64 NumPy / ``math`` exactly as
the reference evaluates them; every method cites the reference lines it follows.  The
Pinocchio-backed systems use ``oracle.robots`` (parity unpinned, see oracle/__init__.py).
"""
import math
import random
import numpy as np

from . import robots


def _softplus_over(alpha, x):
    # log(exp(alpha*x) + 1)/alpha as the reference writes it (environment.py:258-263)
    return math.log(math.exp(alpha * x) + 1) / alpha


class Env:
    """Base: environment.py:10-163."""
    chain = None

    def __init__(self, conf):
        self.conf = conf
        self.nx = conf.nx
        self.nu = conf.na
        self.nq = getattr(conf, 'nq', None)
        self.nv = getattr(conf, 'nv', None)
        self.offset = conf.cost_funct_param[0]           # environment.py:43-44
        self.scale = conf.cost_funct_param[1]
        self.alpha = conf.soft_max_param[0]
        self.alpha2 = conf.soft_max_param[1]
        self.TARGET_STATE = conf.TARGET_STATE

    # environment.py:46-55
    def reset(self):
        c = self.conf
        state = np.zeros(c.nb_state)
        time = random.uniform(c.x_init_min[-1], c.x_init_max[-1])
        for i in range(c.nb_state - 1):
            state[i] = random.uniform(c.x_init_min[i], c.x_init_max[i])
        state[-1] = c.dt * round(time / c.dt)
        return state

    # environment.py:70-78 -- reward at the CURRENT state
    def step(self, weights, state, action):
        return self.simulate(state, action), self.reward(weights, state, action)

    # environment.py:80-91 + robot_utils.py:415-432,399-405 (explicit Euler, old v moves q)
    def simulate(self, state, action):
        nq, nx, dt = self.nq, self.nx, self.conf.dt
        q = np.array(state[:nq], dtype=float)
        v = np.array(state[nq:nx], dtype=float)
        dv = self.chain.forward_dynamics(q, v, np.asarray(action, dtype=float))
        nxt = np.zeros(nx + 1)
        nxt[:nq] = q + v * dt
        nxt[nq:nx] = v + dv * dt
        nxt[-1] = state[-1] + dt
        return nxt

    # environment.py:93-109
    def derivative(self, state, action):
        Minv = self.chain.minv(state[:self.nq])           # = aba_derivatives(...)[2]; the derivative passes are not needed here
        Fu = np.zeros((self.nx + 1, self.nu))
        Fu[self.nv:-1, :] = Minv
        Fu[:self.nx, :] *= self.conf.dt
        return self._normalise_Fu(Fu)

    def _normalise_Fu(self, Fu):
        if self.conf.NORMALIZE_INPUTS:                   # environment.py:106-107
            Fu[:-1] *= (1 / np.asarray(self.conf.state_norm_arr, dtype=float)[:-1, None])
        return Fu

    # environment.py:111-132
    def augmented_derivative(self, state, action):
        nv, nx, dt = self.nv, self.nx, self.conf.dt
        dq, dv, Minv = self.chain.aba_derivatives(state[:self.nq], state[self.nq:nx], action)
        Fx = np.zeros((nx, nx))
        Fu = np.zeros((nx, self.nu))
        Fx[:nv, nv:nx] = np.identity(nv)
        Fx[nv:nx, :nv] = dq
        Fx[nv:nx, nv:nx] = dv
        Fu[nv:nx, :] = Minv
        return np.identity(nx) + dt * Fx, Fu * dt

    # environment.py:134-144 (per-sample loop, output cast to float32)
    def simulate_batch(self, state, action):
        return np.array([self.simulate(s, a) for s, a in zip(state, action)]).astype(np.float32)

    def derivative_batch(self, state, action):
        return np.array([self.derivative(s, a) for s, a in zip(state, action)]).astype(np.float32)

    # environment.py:146-156
    def get_end_effector_position(self, state, recompute=True):
        return np.array(self.chain.ee_position(np.asarray(state[:self.nq], dtype=float)))

    # environment.py:158-163
    def bound_control_cost(self, action):
        u_cost = 0
        for i in range(self.conf.nb_action):
            u_cost += action[i] * action[i] + self.conf.w_b * (action[i] / self.conf.u_max[i]) ** 10
        return u_cost

    # -- shared reward pieces (environment.py:258-275 and the five near-identical twins) --
    def _ellipses(self, p):
        o = self.conf.obs_param
        x, y = p[0], p[1]
        out = []
        for k in range(3):
            xc, yc, a, b = o[2 * k], o[2 * k + 1], o[6 + 2 * k], o[7 + 2 * k]
            e = ((x - xc) ** 2) / ((a / 2) ** 2) + ((y - yc) ** 2) / ((b / 2) ** 2) - 1.0
            out.append(math.log(math.exp(self.alpha * -e) + 1) / self.alpha)
        return out

    def _peak(self, p, dims=2):
        s = 0.0
        for i in range(dims):
            s = s + math.sqrt((p[i] - self.TARGET_STATE[i]) ** 2 + 0.1) - math.sqrt(0.1) - 0.1
        return math.log(math.exp(self.alpha2 * -s) + 1) / self.alpha2

    def _vel_cost(self, weights, state):
        return 0

    def reward(self, weights, state, action=None):
        p = self.get_end_effector_position(state)
        x_ee, y_ee = p[0], p[1]
        ell = self._ellipses(p)
        peak = self._peak(p)
        vel = self._vel_cost(weights, state)
        u_cost = self.bound_control_cost(action) if action is not None else 0
        dist = (x_ee - self.TARGET_STATE[0]) ** 2 + (y_ee - self.TARGET_STATE[1]) ** 2
        return self.scale * (-weights[0] * dist + weights[1] * peak - weights[2] * vel - weights[3] * ell[0]
                             - weights[4] * ell[1] - weights[5] * ell[2] - weights[6] * u_cost + self.offset)

    # environment.py:277-286 etc.: state part per sample in fp64 (no action), cast to f32,
    # action part in float32 "tensor" arithmetic.
    def reward_batch(self, weights, state, action):
        partial = np.array([self.reward(w, s) for w, s in zip(weights, state)])
        action = np.asarray(action, dtype=np.float32)
        w_b = np.float32(self.conf.w_b)
        u_max = np.asarray(self.conf.u_max, dtype=np.float32)
        u_cost = np.sum(action ** 2 + w_b * (action / u_max) ** 10, axis=1, dtype=np.float32)
        r = np.float32(self.scale) * (-(np.asarray(weights)[:, 6].astype(np.float32)) * u_cost) + partial.astype(np.float32)
        return r.reshape(-1, 1).astype(np.float32)

    # dr/da of reward_batch (NeuralNetwork.py:199-204), analytic
    def reward_batch_da(self, weights, action):
        action = np.asarray(action, dtype=np.float64)
        u_max = np.asarray(self.conf.u_max, dtype=np.float64)
        w6 = np.asarray(weights)[:, 6:7]
        return -self.scale * w6 * (2 * action + 10 * self.conf.w_b * action ** 9 / u_max ** 10)


class SingleIntegrator(Env):
    """environment.py:165-286."""

    def simulate(self, state, action):                    # :235-243
        dt = self.conf.dt
        nxt = np.zeros(self.nx + 1)
        nxt[0] = state[0] + dt * action[0]
        nxt[1] = state[1] + dt * action[1]
        nxt[2] = state[2] + dt
        return nxt

    def derivative(self, state, action):                  # :209-219
        Fu = np.zeros((self.nx + 1, self.nu))
        Fu[0, 0] = self.conf.dt
        Fu[1, 1] = self.conf.dt
        return self._normalise_Fu(Fu)

    def augmented_derivative(self, state, action):        # :221-233
        Fx = np.identity(2)
        Fu = np.zeros((2, 2))
        Fu[0, 0] = self.conf.dt
        Fu[1, 1] = self.conf.dt
        return Fx, Fu

    def get_end_effector_position(self, state, recompute=True):   # :245-250
        p = np.zeros(3)
        p[:2] = state[:2]
        return p


class DoubleIntegrator(Env):
    """environment.py:288-362; dynamics through Pinocchio on double_integrator.urdf."""
    chain = robots.DOUBLE_INTEGRATOR


class Car(Env):
    """environment.py:364-491."""

    def simulate(self, state, action):                    # :437-448
        dt = self.conf.dt
        c, s = math.cos(state[2]), math.sin(state[2])
        nxt = np.zeros(self.nx + 1)
        nxt[0] = state[0] + dt * state[3] * c + dt ** 2 * state[4] * c / 2
        nxt[1] = state[1] + dt * state[3] * s + dt ** 2 * state[4] * s / 2
        nxt[2] = state[2] + dt * action[0]
        nxt[3] = state[3] + dt * state[4]
        nxt[4] = state[4] + dt * action[1]
        nxt[5] = state[5] + dt
        return nxt

    def derivative(self, state, action):                  # :408-418
        Fu = np.zeros((self.nx + 1, self.nu))
        Fu[2, 0] = self.conf.dt
        Fu[4, 1] = self.conf.dt
        return self._normalise_Fu(Fu)

    def augmented_derivative(self, state, action):        # :420-435
        dt = self.conf.dt
        c, s = math.cos(state[2]), math.sin(state[2])
        Fx = np.array([[1, 0, -dt * state[3] * s - dt ** 2 * state[4] * s / 2, dt * c, dt ** 2 * c / 2],
                       [0, 1, dt * state[3] * c + dt ** 2 * state[4] * c / 2, dt * s, dt ** 2 * s / 2],
                       [0, 0, 1, 0, 0],
                       [0, 0, 0, 1, dt],
                       [0, 0, 0, 0, 1]], dtype=float)
        Fu = np.zeros((5, 2))
        Fu[2, 0] = dt
        Fu[4, 1] = dt
        return Fx, Fu

    def get_end_effector_position(self, state, recompute=True):   # :450-455
        p = np.zeros(3)
        p[:2] = state[:2]
        return p


class CarPark(Car):
    """environment.py:493-652."""

    def simulate(self, state, action):                    # :584-595
        c_ = self.conf
        dt = c_.dt
        nxt = np.zeros(self.nx + 1)
        nxt[0] = state[0] + dt * state[3] * math.cos(state[2])
        nxt[1] = state[1] + dt * state[3] * math.sin(state[2])
        nxt[2] = state[2] + dt * state[3] * math.tan(state[4]) / c_.L_delta
        nxt[3] = state[3] + dt * action[0]
        nxt[4] = state[4] + dt * action[1] / c_.tau_delta
        nxt[5] = state[5] + dt
        return nxt

    def derivative(self, state, action):                  # :555-565
        Fu = np.zeros((self.nx + 1, self.nu))
        Fu[3, 0] = self.conf.dt
        Fu[4, 1] = self.conf.dt / self.conf.tau_delta
        return self._normalise_Fu(Fu)

    def augmented_derivative(self, state, action):        # :567-582 (mpmath.sec -> 1/cos)
        c_ = self.conf
        dt = c_.dt
        c, s = math.cos(state[2]), math.sin(state[2])
        sec2 = (1.0 / math.cos(state[4])) ** 2
        Fx = np.array([[1, 0, -dt * state[3] * s, dt * c, 0],
                       [0, 1, dt * state[3] * c, dt * s, 0],
                       [0, 0, 1, dt * math.tan(state[4]) / c_.L_delta, dt * state[3] * sec2 / c_.L_delta],
                       [0, 0, 0, 1, 0],
                       [0, 0, 0, 0, 1]], dtype=float)
        Fu = np.zeros((5, 2))
        Fu[3, 0] = dt
        Fu[4, 1] = dt / c_.tau_delta
        return Fx, Fu

    def get_end_effector_position(self, state, recompute=True):   # :597-602
        p = np.zeros(3)
        th = state[2]
        Rm = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        p[:2] = state[:2] + Rm.dot(np.array([self.conf.L_delta / 2, 0]))
        return p

    def obs_cost_fun(self, x, y, x_step, y_step, Wx, Wy, fv=1, k=50):     # :604-613
        k = self.conf.k_db
        term1 = 4 + 4 * (y - y_step + Wy / 2) ** 2 * k ** 2
        term2 = 4 + 4 * (y - y_step - Wy / 2) ** 2 * k ** 2
        term3 = 4 + 4 * (x - x_step + Wx / 2) ** 2 * k ** 2
        term4 = 4 + 4 * (x - x_step - Wx / 2) ** 2 * k ** 2
        return ((term1) ** (-1 / 2) * fv * (-np.sqrt(term2) / 2 + (y - y_step - Wy / 2) * k) * (term3) ** (-1 / 2)
                * (term2) ** (-1 / 2) * (np.sqrt(term1) / 2 + (y - y_step + Wy / 2) * k) * (term4) ** (-1 / 2)
                * (np.sqrt(term3) / 2 + (x - x_step + Wx / 2) * k) * (-np.sqrt(term4) / 2 + (x - x_step - Wx / 2) * k))

    def reward(self, weights, state, action=None):        # :615-641
        c_ = self.conf
        o = c_.obs_param
        p = self.get_end_effector_position(state)
        x_ee, y_ee = p[0], p[1]
        th = state[2]
        Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        cp = np.dot(Rm, np.asarray(c_.check_points_BF).T).T + np.array([x_ee, y_ee])
        obs_cost = 0
        for k in range(3):
            obs_cost += np.sum(self.obs_cost_fun(cp[:, 0], cp[:, 1], o[2 * k], o[2 * k + 1], o[6 + 2 * k], o[7 + 2 * k]))
        peak = self._peak(p)
        u_cost = self.bound_control_cost(action) if action is not None else 0
        dist = (x_ee - self.TARGET_STATE[0]) ** 2 + (y_ee - self.TARGET_STATE[1]) ** 2
        return self.scale * (-weights[0] * dist + weights[1] * peak - weights[2] * state[3] ** 2
                             - weights[3] * obs_cost - weights[6] * u_cost + self.offset)


class Manipulator(Env):
    """environment.py:654-734."""
    chain = robots.MANIPULATOR

    def _vel_cost(self, weights, state):                  # :709-712
        if weights[2] != 0:
            v = np.asarray(state[self.nq:self.nx], dtype=float)
            return v.dot(v)
        return 0


class UR5(Env):
    """environment.py:736-816."""
    chain = robots.UR5

    def reward(self, weights, state, action=None):        # :780-805
        o = self.conf.obs_param
        p = self.get_end_effector_position(state)
        ell = []
        for k in range(3):
            xc, yc, zc = o[3 * k], o[3 * k + 1], o[3 * k + 2]
            a, b, c = o[9 + 3 * k], o[10 + 3 * k], o[11 + 3 * k]
            e = (((p[0] - xc) ** 2) / ((a / 2) ** 2) + ((p[1] - yc) ** 2) / ((b / 2) ** 2)
                 + ((p[2] - zc) ** 2) / ((c / 2) ** 2) - 1.0)
            ell.append(math.log(math.exp(self.alpha * -e) + 1) / self.alpha)
        peak = self._peak(p, dims=3)
        if action is not None:
            action = np.asarray(action, dtype=float)
            u_cost = action.dot(action)                   # quirk Q8
        else:
            u_cost = 0
        v = np.asarray(state[self.nq:self.nx], dtype=float)
        vel = v.dot(v)
        T = self.TARGET_STATE
        dist = (p[0] - T[0]) ** 2 + (p[1] - T[1]) ** 2 + (p[2] - T[2]) ** 2
        return self.scale * (-weights[0] * dist + weights[1] * peak - weights[2] * vel - weights[3] * ell[0]
                             - weights[4] * ell[1] - weights[5] * ell[2] - weights[6] * u_cost + self.offset)


SYSTEMS = {'single_integrator': SingleIntegrator, 'double_integrator': DoubleIntegrator, 'car': Car,
           'car_park': CarPark, 'manipulator': Manipulator, 'ur5': UR5}


def make_env(conf):
    return SYSTEMS[conf.system_id](conf)


# ------------------------------------------------------------------------------------------------------------------
# The per-sample loops of simulate_batch / derivative_batch (environment.py:134-144) fanned over forked worker processes:
# same arithmetic per sample, only the wall time changes (NumPy RNEA: ~10 ms per manipulator sample, ~40 ms per UR5 sample).
# Used by the large-batch parity tests (B = 4096 / 16384) and by bench.py's CPU baseline, never by the product.
_POOL_ENV = None


def _pool_sim(args):
    s, a = args
    return np.array([_POOL_ENV.simulate(x, u) for x, u in zip(s, a)])


def _pool_der(args):
    s, a = args
    return np.array([_POOL_ENV.derivative(x, u) for x, u in zip(s, a)])


class PooledEnv:
    """Wraps an oracle Env: simulate_batch / derivative_batch split the batch into chunks over ``procs`` forked workers;
    every other attribute is the wrapped environment's."""

    def __init__(self, env, procs=None):
        import os
        self._env = env
        self._procs = int(procs or min(32, os.cpu_count() or 1))

    def __getattr__(self, name):
        return getattr(self._env, name)

    def _map(self, fn, state, action):
        import multiprocessing as mp
        global _POOL_ENV
        state, action = np.asarray(state), np.asarray(action)
        B = len(state)
        if self._procs <= 1 or B < 4 * self._procs:
            _POOL_ENV = self._env
            return fn((state, action))
        _POOL_ENV = self._env                      # inherited by the forked workers
        n = self._procs * 4
        cuts = [B * i // n for i in range(n + 1)]
        with mp.get_context('fork').Pool(self._procs) as pool:
            parts = pool.map(fn, [(state[a:b], action[a:b]) for a, b in zip(cuts[:-1], cuts[1:]) if b > a])
        return np.concatenate(parts, axis=0)

    def simulate_batch(self, state, action):
        return self._map(_pool_sim, state, action).astype(np.float32)

    def derivative_batch(self, state, action):
        return self._map(_pool_der, state, action).astype(np.float32)
