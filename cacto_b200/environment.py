"""Host-side mirror of the reference's ``environment.py`` for the GPU hot path.

Same class names, method names, argument order and return shapes as the reference
(environment.py:10-816), but every numeric method launches a kernel of libcacto_b200.so through
the C ABI (include/cacto_b200.h) instead of looping over samples in Python / Pinocchio:

  simulate, simulate_batch            -> cacto_dyn_step          (environment.py:80-91,134-138)
  derivative, derivative_batch        -> cacto_dyn_derivative    (environment.py:93-109,140-144)
  augmented_derivative[_batch]        -> cacto_dyn_augmented     (environment.py:111-132; TO.py:181)
  get_end_effector_position[_batch]   -> cacto_ee_position       (environment.py:146-156)
  reward, reward_batch[_da]           -> cacto_reward            (environment.py:252-286 and twins)

Single-sample methods take/return NumPy fp64 like the reference; the ``*_batch`` methods take NumPy
arrays or torch tensors and return CUDA tensors (float32 unless given float64), the analogue of the
reference's ``tf.convert_to_tensor(..., float32)``.  There is no CPU fallback: without a CUDA device
or without the shared library these classes raise.
"""
import ctypes as C
import math
import random

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from .ops import ops, sys_tensor

_DT = {torch.float32: 0, torch.float64: 1}


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('cacto_b200 needs a CUDA device (no CPU fallback on the hot path)')
    return torch.device('cuda', torch.cuda.current_device())


def _as_cuda(x, dtype=None):
    """NumPy / torch -> contiguous CUDA tensor; dtype None keeps float64/float32, maps others to f32."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is None:
        dtype = t.dtype if t.dtype in (torch.float32, torch.float64) else torch.float32
    return t.to(device=_device(), dtype=dtype).contiguous()


class Env:
    """environment.py:10-163."""

    def __init__(self, conf):
        self.conf = conf
        self.nq = getattr(conf, 'nq', None)
        self.nv = getattr(conf, 'nv', None)
        self.nx = conf.nx
        self.nu = conf.na
        self.offset = conf.cost_funct_param[0]
        self.scale = conf.cost_funct_param[1]
        self.alpha = conf.soft_max_param[0]
        self.alpha2 = conf.soft_max_param[1]
        self.TARGET_STATE = conf.TARGET_STATE
        o = conf.obs_param
        if len(o) == 12:
            (self.XC1, self.YC1, self.XC2, self.YC2, self.XC3, self.YC3,
             self.A1, self.B1, self.A2, self.B2, self.A3, self.B3) = [o[i] for i in range(12)]
        self.params = _lib.make_sys_params(conf)
        self._p = C.byref(self.params)                 # for the entry points still bound with ctypes
        self._pt = sys_tensor(self.params)             # for the torch.ops.cacto.* custom ops
        self.ns = int(conf.nb_state)
        self.na = int(conf.nb_action)

    # ------------------------------------------------------------------ host-side pieces
    def reset(self):
        """environment.py:46-55 (host RNG stream kept: random.uniform)."""
        c = self.conf
        state = np.zeros(c.nb_state)
        time = random.uniform(c.x_init_min[-1], c.x_init_max[-1])
        for i in range(c.nb_state - 1):
            state[i] = random.uniform(c.x_init_min[i], c.x_init_max[i])
        state[-1] = c.dt * round(time / c.dt)
        return state

    def check_ICS_feasible(self, state):
        """environment.py:57-68."""
        p = self.get_end_effector_position(state)
        e = [((p[0] - xc) ** 2) / ((a / 2) ** 2) + ((p[1] - yc) ** 2) / ((b / 2) ** 2)
             for xc, yc, a, b in ((self.XC1, self.YC1, self.A1, self.B1), (self.XC2, self.YC2, self.A2, self.B2),
                                  (self.XC3, self.YC3, self.A3, self.B3))]
        return e[0] > 1 and e[1] > 1 and e[2] > 1

    def bound_control_cost(self, action):
        """environment.py:158-163."""
        u_cost = 0
        for i in range(self.conf.nb_action):
            u_cost += action[i] * action[i] + self.conf.w_b * (action[i] / self.conf.u_max[i]) ** 10
        return u_cost

    def step(self, weights, state, action):
        """environment.py:70-78: (simulate(s, a), reward(w, s, a)) -- reward at the CURRENT state."""
        return self.simulate(state, action), self.reward(weights, state, action)

    # ------------------------------------------------------------------ batched kernels
    def _sa(self, state, action, dtype=None):
        s = _as_cuda(state, dtype)
        a = _as_cuda(action, s.dtype)
        if s.dim() != 2 or s.shape[1] != self.ns or a.shape != (s.shape[0], self.na):
            raise ValueError(f'expected state [B,{self.ns}] and action [B,{self.na}], got {tuple(s.shape)} {tuple(a.shape)}')
        return s, a

    def simulate_batch(self, state, action):
        """environment.py:134-138 -> [B, ns]."""
        s, a = self._sa(state, action)
        out = torch.empty_like(s)
        ops.dyn_step(self._pt, 0, s, a, out)
        return out

    def derivative_batch(self, state, action):
        """environment.py:140-144 -> [B, ns, na] (normalised, zero time row)."""
        s, a = self._sa(state, action)
        out = torch.empty((s.shape[0], self.ns, self.na), dtype=s.dtype, device=s.device)
        ops.dyn_derivative(self._pt, 0, s, a, out)
        return out

    def augmented_derivative_batch(self, state, action):
        """Batched environment.py:111-132 -> (Fx[B, nx, nx], Fu[B, nx, na]); state may be [B, nx] or [B, ns]."""
        s = _as_cuda(state)
        if s.shape[1] == self.nx:
            s = torch.cat([s, torch.zeros((s.shape[0], 1), dtype=s.dtype, device=s.device)], dim=1)
        s, a = self._sa(s, action)
        B = s.shape[0]
        Fx = torch.empty((B, self.nx, self.nx), dtype=s.dtype, device=s.device)
        Fu = torch.empty((B, self.nx, self.na), dtype=s.dtype, device=s.device)
        ops.dyn_augmented(self._pt, 0, s, a, Fx, Fu)
        return Fx, Fu

    def get_end_effector_position_batch(self, state):
        s = _as_cuda(state)
        out = torch.empty((s.shape[0], 3), dtype=s.dtype, device=s.device)
        ops.ee_position(self._pt, 0, s.contiguous(), out)
        return out

    def _weights8(self, weights, B):
        w = torch.as_tensor(np.asarray(weights, dtype=np.float64)).reshape(B, -1)
        w8 = torch.zeros((B, 8), dtype=torch.float64)
        w8[:, :w.shape[1]] = w
        return w8.to(_device())

    def _reward_kernel(self, weights, state, action, plain_ucost, want_grad):
        s = _as_cuda(state)
        B = s.shape[0]
        a = _as_cuda(action, s.dtype) if action is not None else None
        w8 = weights if (isinstance(weights, torch.Tensor) and weights.is_cuda and weights.shape == (B, 8)
                         and weights.dtype == torch.float64) else self._weights8(weights, B)
        r = torch.empty((B,), dtype=s.dtype, device=s.device)
        g = torch.empty((B, self.na), dtype=s.dtype, device=s.device) if want_grad else None
        ops.reward(self._pt, 0, w8.contiguous(), s, a, int(plain_ucost), r, g)
        return r, g

    def reward_batch(self, weights, state, action):
        """environment.py:277-286 (and twins) -> [B, 1]."""
        r, _ = self._reward_kernel(weights, state, action, False, False)
        return r.reshape(-1, 1)

    def reward_batch_da(self, weights, state, action):
        """d reward_batch / d action, the quantity NeuralNetwork.py:199-204 obtains from a GradientTape."""
        _, g = self._reward_kernel(weights, state, action, False, True)
        return g

    # ------------------------------------------------------------------ single-sample API (NumPy fp64)
    def simulate(self, state, action):
        """environment.py:80-91 (and :235, :437, :584)."""
        s = np.asarray(state, dtype=np.float64).reshape(1, -1)
        a = np.asarray(action, dtype=np.float64).reshape(1, -1)
        return self.simulate_batch(s, a)[0].cpu().numpy()

    def derivative(self, state, action):
        """environment.py:93-109 -> Fu[ns, na]."""
        s = np.asarray(state, dtype=np.float64).reshape(1, -1)
        a = np.asarray(action, dtype=np.float64).reshape(1, -1)
        return self.derivative_batch(s, a)[0].cpu().numpy()

    def augmented_derivative(self, state, action):
        """environment.py:111-132 -> (Fx[nx, nx], Fu[nx, na]); ``state`` has nx (TO.py:181) or ns entries."""
        s = np.asarray(state, dtype=np.float64).reshape(1, -1)
        a = np.asarray(action, dtype=np.float64).reshape(1, -1)
        Fx, Fu = self.augmented_derivative_batch(s, a)
        return Fx[0].cpu().numpy(), Fu[0].cpu().numpy()

    def get_end_effector_position(self, state, recompute=True):
        """environment.py:146-156."""
        s = np.zeros((1, self.ns))
        st = np.asarray(state, dtype=np.float64).reshape(-1)
        s[0, :min(len(st), self.ns)] = st[:self.ns]
        return self.get_end_effector_position_batch(s)[0].cpu().numpy()

    _plain_ucost_in_reward = False

    def reward(self, weights, state, action=None):
        """environment.py:252-275 and twins -> float."""
        s = np.asarray(state, dtype=np.float64).reshape(1, -1)
        a = None if action is None else np.asarray(action, dtype=np.float64).reshape(1, -1)
        r, _ = self._reward_kernel(np.asarray(weights, dtype=np.float64).reshape(1, -1), s, a, self._plain_ucost_in_reward, False)
        return float(r[0])


class SingleIntegrator(Env):
    """environment.py:165-286."""


class DoubleIntegrator(Env):
    """environment.py:288-362."""


class Car(Env):
    """environment.py:364-491."""


class CarPark(Car):
    """environment.py:493-652."""

    def check_ICS_feasible(self, state):
        """environment.py:537-553 (host; not on the hot path)."""
        c = self.conf
        p = self.get_end_effector_position(state)
        th = state[2]
        Rm = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        ok = True
        for cp in np.asarray(c.check_points_BF):
            w = np.array(p[:2]) + Rm.dot(cp)
            vals = [self.obs_cost_fun(w[0], w[1], xc, yc, a, b) for xc, yc, a, b in
                    ((self.XC1, self.YC1, self.A1, self.B1), (self.XC2, self.YC2, self.A2, self.B2),
                     (self.XC3, self.YC3, self.A3, self.B3))]
            ok = all(v < 0.5 for v in vals)
            if not ok:
                return ok
        return ok

    def obs_cost_fun(self, x, y, x_step, y_step, Wx, Wy, fv=1, k=50):
        """environment.py:604-613 as a product of four smooth steps (host helper)."""
        k = self.conf.k_db
        sg = lambda z: 0.5 * (1 + z / np.sqrt(1 + z * z))
        return fv * (sg((y - y_step + Wy / 2) * k) * (1 - sg((y - y_step - Wy / 2) * k))
                     * sg((x - x_step + Wx / 2) * k) * (1 - sg((x - x_step - Wx / 2) * k)))


class Manipulator(Env):
    """environment.py:654-734."""


class UR5(Env):
    """environment.py:736-816.  UR5.reward uses u.u as control cost while reward_batch uses the bounded
    cost -- both reproduced."""
    _plain_ucost_in_reward = True


ENVIRONMENTS = dict(single_integrator=SingleIntegrator, double_integrator=DoubleIntegrator, car=Car, car_park=CarPark,
                    manipulator=Manipulator, ur5=UR5)


def make_env(conf):
    return ENVIRONMENTS[conf.system_id](conf)
