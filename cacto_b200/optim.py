"""tf.keras.optimizers.Adam (TF 2.11) and PiecewiseConstantDecay on the fused CUDA Adam kernel.

Reference use: RL.py:82-88 (construction, optional schedule), :105,:109 (apply_gradients).
TF-2.11 update: t = iterations + 1; alpha_t = lr(iterations) sqrt(1 - b2^t) / (1 - b1^t);
m += (g - m)(1 - b1); v += (g^2 - v)(1 - b2); p -= alpha_t m / (sqrt(v) + eps), eps = 1e-7 un-corrected.

alpha_t is evaluated on the device (``cacto_adam_schedule``: step counter, schedule table and alpha live in
HBM) so that a captured CUDA graph of the whole update replays without host-side scalars; ``iterations`` on
the host mirrors the device counter.
"""
import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr
from .NeuralNetwork import Network
from .ops import ops


class PiecewiseConstantDecay:
    """values[#{boundaries < step}]  (tf.keras.optimizers.schedules.PiecewiseConstantDecay)."""

    def __init__(self, boundaries, values):
        assert len(values) == len(boundaries) + 1
        self.boundaries, self.values = list(boundaries), list(values)

    def __call__(self, step):
        return self.values[sum(1 for b in self.boundaries if b < step)]


class Adam:
    def __init__(self, learning_rate, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self._state = {}
        self._dev = None

    def current_lr(self):
        return self.learning_rate(self.iterations) if callable(self.learning_rate) else self.learning_rate

    def _device_state(self, device):
        if self._dev is None:
            if isinstance(self.learning_rate, PiecewiseConstantDecay):
                b, v = self.learning_rate.boundaries, self.learning_rate.values
            elif callable(self.learning_rate):
                raise TypeError('only float or PiecewiseConstantDecay learning rates are supported')
            else:
                b, v = [], [self.learning_rate]
            self._dev = dict(step=torch.full((1,), self.iterations, dtype=torch.int64, device=device),
                             alpha=torch.zeros(1, dtype=torch.float32, device=device),
                             boundaries=torch.tensor(b if b else [0.0], dtype=torch.float32, device=device),
                             values=torch.tensor(v, dtype=torch.float32, device=device), nb=len(b))
        return self._dev

    def moments(self, net):
        st = self._state.get(id(net))
        if st is None:
            st = (torch.zeros_like(net.params), torch.zeros_like(net.params))
            self._state[id(net)] = st
        return st

    def prepare(self, device, zero=None):
        """Launch the schedule kernel: alpha_t for the coming step, step counter += 1, ``zero[0] = 0``
        (an optional one-element float tensor, e.g. the loss accumulator the next kernel adds into)."""
        d = self._device_state(device)
        ops.adam_schedule(d['step'], d['boundaries'], d['values'], d['nb'], self.beta_1, self.beta_2, d['alpha'], zero)

    @staticmethod
    def prepare_pair(first, second, device, zero=None):
        """``first.prepare(device, zero)`` and ``second.prepare(device)`` in ONE launch (the two optimizers of RL_AC.update)."""
        a, b = first._device_state(device), second._device_state(device)
        ops.adam_schedule2(a['step'], a['boundaries'], a['values'], a['nb'], first.beta_1, first.beta_2, a['alpha'],
                           b['step'], b['boundaries'], b['values'], b['nb'], second.beta_1, second.beta_2, b['alpha'], zero)

    def step(self, net, target=None, tau=0.0, prepared=False, peer=None, zero_other=None):
        """One Adam step on ``net`` from ``net.grad`` (zeroed afterwards); optionally the Polyak update
        ``target = tau * net + (1 - tau) * target`` (RL.py:113-118) in the same launch.

        ``peer`` = ``PeerReduce.table(net)`` switches to the data-parallel kernel: the gradient blocks of all ranks are summed
        over NVLink inside the launch (no all-reduce before it); ``net.grad`` is then left as is and ``zero_other`` -- the
        gradient block of the network stepped before this one -- is cleared instead (include/cacto_b200.h)."""
        m, v = self.moments(net)
        if not prepared:
            self.prepare(net.params.device)
        d = self._device_state(net.params.device)
        tgt = ptr(target.params if target is not None else None)
        if peer is not None:
            grads, flags, rank, max_ctas = peer
            check(lib.cacto_adam_step_peer(ptr(net.params), grads, flags, len(grads), rank, ptr(zero_other),
                                           zero_other.numel() if zero_other is not None else 0, ptr(m), ptr(v), ptr(d['alpha']), self.beta_1,
                                           self.beta_2, self.epsilon, tgt, float(tau), ptr(net.params_T), net.is_critic, net.ns, net.na, net.n,
                                           max_ctas, stream_ptr()), 'adam_step_peer')
        else:
            ops.adam_step(net.params, net.grad, m, v, 0.0, d['alpha'], self.beta_1, self.beta_2, self.epsilon, target.params if target is not None else None,
                          float(tau), net.params_T, net.is_critic, net.ns, net.na)
        self.iterations += 1

    def apply_gradients(self, grads_and_vars, prepared=False):
        """Keras signature: ``apply_gradients(zip(grads, model.trainable_variables))``.  ``prepared``: the schedule kernel of this
        step has been launched already (``prepare`` / ``prepare_pair``)."""
        grads, variables = zip(*list(grads_and_vars))
        net = Network._registry.get(variables[0].data_ptr())
        if net is None:
            raise ValueError('variables do not belong to a cacto_b200 Network')
        if grads[0].data_ptr() != net.grad.data_ptr():          # foreign gradients: stage them into the accumulator
            for dst, g in zip(net._grad_views, grads):
                dst.copy_(torch.as_tensor(g).to(dst.device, dst.dtype))
        self.step(net, prepared=prepared)
