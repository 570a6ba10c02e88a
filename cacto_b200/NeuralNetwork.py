"""Host-side mirror of the reference's ``NeuralNetwork.py`` (class ``NN``) on the CUDA kernels.

Networks are flat float32 parameter blocks in HBM in Keras order [W1 (in x out), b1, W2, b2, ...]
(include/cacto_b200.h), wrapped by ``Network`` which offers the slice of the Keras model surface the
reference uses: ``trainable_variables`` / ``variables``, ``get_weights`` / ``set_weights``,
``save_weights`` / ``load_weights`` and ``__call__``.

  NN.eval                 -> cacto_actor_forward / cacto_critic_forward   (NeuralNetwork.py:130-138)
  NN.compute_critic_grad  -> cacto_critic_grad                            (NeuralNetwork.py:150-178)
  NN.compute_actor_grad   -> cacto_actor_grad                             (NeuralNetwork.py:180-232)

The regularisers the reference attaches to its layers are inert there (gradients are taken of the loss
only, ``model.losses`` is never added: SURVEY.md quirk Q6) and are therefore not represented.
Only the default critic ('sine', every conf) is on the hot path; the elu / relu / sine-elu variants are
listed as next work in DESIGN.md and raise here.
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from .environment import _as_cuda, _device

CRITIC_HIDDEN = (64, 64, 128, 128)


class Network:
    """A dense network stored as one flat float32 block (+ its per-layer transposed copy and a gradient
    accumulator).  kind: 'actor' (LeakyReLU 0.3) or 'critic_sine'."""

    _registry = {}

    def __init__(self, kind, ns, na, dims):
        self.kind, self.ns, self.na, self.dims = kind, int(ns), int(na), list(dims)
        self.n = sum(i * o + o for i, o in zip(dims[:-1], dims[1:]))
        expect = lib.cacto_critic_param_count(self.ns) if kind == 'critic_sine' else lib.cacto_actor_param_count(self.ns, self.na)
        assert self.n == expect, (self.n, expect)
        dev = _device()
        self.params = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.params_T = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self._views = self._make_views(self.params)
        self._grad_views = self._make_views(self.grad)
        Network._registry[self.params.data_ptr()] = self

    @property
    def is_critic(self):
        return int(self.kind == 'critic_sine')

    def _make_views(self, flat):
        out, o = [], 0
        for i, k in zip(self.dims[:-1], self.dims[1:]):
            out.append(flat[o:o + i * k].view(i, k))
            o += i * k
            out.append(flat[o:o + k])
            o += k
        return out

    # -- Keras-like surface -------------------------------------------------------------------
    @property
    def trainable_variables(self):
        return self._views

    variables = trainable_variables

    def get_weights(self):
        return [v.detach().cpu().numpy().copy() for v in self._views]

    def set_weights(self, weights):
        assert len(weights) == len(self._views)
        for v, w in zip(self._views, weights):
            w = torch.as_tensor(np.asarray(w, dtype=np.float32))
            assert tuple(w.shape) == tuple(v.shape), (tuple(w.shape), tuple(v.shape))
            v.copy_(w)
        self.refresh_transposed()

    def refresh_transposed(self):
        check(lib.cacto_transpose_params(ptr(self.params), ptr(self.params_T), self.is_critic, self.ns, self.na, stream_ptr()),
              'transpose_params')

    def save_weights(self, path):
        """Checkpoint as .npz (the reference writes Keras .h5, RL.py:191-195; .h5 export is listed as next work)."""
        np.savez(path if str(path).endswith('.npz') else str(path) + '.npz', *self.get_weights())

    def load_weights(self, path):
        """``.npz`` written by save_weights, or a Keras ``.h5`` file written by the reference (RL.py:91-97,191-195)."""
        path = str(path)
        if path.endswith('.h5'):
            from .h5weights import load_keras_weights
            self.set_weights(load_keras_weights(path, self.ns))
            return
        z = np.load(path if path.endswith('.npz') else path + '.npz')
        self.set_weights([z[f'arr_{i}'] for i in range(len(self._views))])

    def __call__(self, x, training=True):
        raise TypeError('call NN.eval(network, input): normalisation is fused into the forward kernel')


def _glorot(rng, i, o):
    lim = math.sqrt(6.0 / (i + o))
    return rng.uniform(-lim, lim, (i, o)).astype(np.float32)


class NN:
    """NeuralNetwork.py:10-232."""

    def __init__(self, env, conf, w_S=0, seed=None):
        self.env = env
        self.conf = conf
        self.w_S = w_S
        self._rng = np.random.default_rng(seed)
        self._p = env._p
        self.last_critic_loss = torch.zeros(1, dtype=torch.float32, device=_device())

    # -- model constructors ---------------------------------------------------------------------
    def create_actor(self):
        """NeuralNetwork.py:51-63: Dense(NH1) LeakyReLU Dense(NH2) LeakyReLU Dense(na); glorot-uniform kernels, zero biases."""
        c = self.conf
        if (c.NH1, c.NH2) != (256, 256):
            raise NotImplementedError('the fused actor kernels are specialised for NH1 = NH2 = 256 (every reference conf)')
        net = Network('actor', c.nb_state, c.nb_action, [c.nb_state, c.NH1, c.NH2, c.nb_action])
        w = []
        for i, o in zip(net.dims[:-1], net.dims[1:]):
            w += [_glorot(self._rng, i, o), np.zeros(o, np.float32)]
        net.set_weights(w)
        return net

    def create_critic_sine(self):
        """NeuralNetwork.py:95-108: four SIREN layers (64, 64, 128, 128; tf_siren w0 = 1: kernel U(+-sqrt(6/fan_in)),
        bias he_uniform U(+-sqrt(6/units))) and a linear Dense(1)."""
        c = self.conf
        net = Network('critic_sine', c.nb_state, c.nb_action, [c.nb_state] + list(CRITIC_HIDDEN) + [1])
        w = []
        for i, o in zip(net.dims[:-2], net.dims[1:-1]):
            lk, lb = math.sqrt(6.0 / i), math.sqrt(6.0 / o)
            w += [self._rng.uniform(-lk, lk, (i, o)).astype(np.float32), self._rng.uniform(-lb, lb, o).astype(np.float32)]
        w += [_glorot(self._rng, net.dims[-2], 1), np.zeros(1, np.float32)]
        net.set_weights(w)
        return net

    def create_critic_elu(self):
        raise NotImplementedError("critic_type 'elu' is not on the GPU hot path yet (every reference conf uses 'sine')")

    create_critic_sine_elu = create_critic_relu = create_critic_elu

    # -- forward ------------------------------------------------------------------------------
    def eval(self, NN, input):
        """NeuralNetwork.py:130-138: normalise (utils.py:17-24) + forward.  -> [B, na] or [B, 1] float32."""
        x = _as_cuda(input, torch.float32)
        if x.dim() == 1:
            x = x.reshape(1, -1)
        B = x.shape[0]
        if NN.kind == 'actor':
            out = torch.empty((B, NN.na), dtype=torch.float32, device=x.device)
            check(lib.cacto_actor_forward(self._p, ptr(NN.params), ptr(x), ptr(out), B, stream_ptr()), 'actor_forward')
        else:
            out = torch.empty((B, 1), dtype=torch.float32, device=x.device)
            check(lib.cacto_critic_forward(self._p, ptr(NN.params), ptr(x), ptr(out), ptr(None), B, stream_ptr()), 'critic_forward')
        return out

    def eval_with_gradient(self, critic, input):
        """(V, dV/ds) with dV/ds taken w.r.t. the RAW state as the reference's tapes do (NeuralNetwork.py:162-165,190-195)."""
        x = _as_cuda(input, torch.float32)
        B = x.shape[0]
        V = torch.empty((B, 1), dtype=torch.float32, device=x.device)
        dV = torch.empty((B, critic.ns), dtype=torch.float32, device=x.device)
        check(lib.cacto_critic_forward(self._p, ptr(critic.params), ptr(x), ptr(V), ptr(dV), B, stream_ptr()), 'critic_forward')
        return V, dV

    def custom_logarithm(self, input):
        """NeuralNetwork.py:140-148."""
        x = torch.as_tensor(input)
        pos = torch.log(torch.clamp(x, min=1e-7) + 1)
        neg = -torch.log(torch.clamp(-x, min=1e-7) + 1)
        return torch.where(x > 0, pos, neg)

    # -- gradients ----------------------------------------------------------------------------
    def compute_critic_grad(self, critic_model, target_critic, state_batch, state_next_rollout_batch, partial_reward_to_go_batch,
                            dVdx_batch, d_batch, weights_batch, global_batch=None):
        """NeuralNetwork.py:150-178 -> (critic_grad, reward_to_go, critic_value, target_critic_value(state)).
        ``global_batch`` is the batch size the loss is averaged over (data-parallel shards pass the global one)."""
        f32 = torch.float32
        s = _as_cuda(state_batch, f32)
        B = s.shape[0]
        sn = _as_cuda(state_next_rollout_batch, f32)
        pr = _as_cuda(partial_reward_to_go_batch, f32).reshape(-1)
        dv = _as_cuda(dVdx_batch, f32)
        d = _as_cuda(d_batch, f32).reshape(-1)
        w = _as_cuda(weights_batch, f32).reshape(-1)
        dev = s.device
        rtg = torch.empty((B, 1), dtype=f32, device=dev)
        V = torch.empty((B, 1), dtype=f32, device=dev)
        Vt = torch.empty((B, 1), dtype=f32, device=dev)
        critic_model.grad.zero_()
        self.last_critic_loss.zero_()
        inv_B = 1.0 / float(global_batch if global_batch is not None else B)
        check(lib.cacto_critic_grad(self._p, ptr(critic_model.params), ptr(critic_model.params_T), ptr(target_critic.params),
                                    float(self.w_S), int(bool(self.conf.MC)), ptr(s), ptr(sn), ptr(pr), ptr(dv), ptr(d), ptr(w), inv_B,
                                    ptr(critic_model.grad), ptr(rtg), ptr(V), ptr(Vt), ptr(self.last_critic_loss), B, stream_ptr()),
              'critic_grad')
        return critic_model._grad_views, rtg, V, Vt

    def compute_actor_grad(self, actor_model, critic_model, state_batch, term_batch, batch_size, global_batch=None, return_actions=False):
        """NeuralNetwork.py:180-232 -> actor_grad (list of per-variable views of the gradient block)."""
        s = _as_cuda(state_batch, torch.float32)
        B = s.shape[0]
        if batch_size is None:
            batch_size = self.conf.BATCH_SIZE
        term = _as_cuda(term_batch, torch.float64).reshape(-1)
        assert term.numel() == B
        actions = torch.empty((B, actor_model.na), dtype=torch.float32, device=s.device) if return_actions else None
        actor_model.grad.zero_()
        inv_B = 1.0 / float(global_batch if global_batch is not None else B)
        check(lib.cacto_actor_grad(self._p, ptr(actor_model.params), ptr(actor_model.params_T), ptr(critic_model.params),
                                   ptr(critic_model.params_T), ptr(s), ptr(term), inv_B, ptr(actor_model.grad), ptr(actions), B,
                                   stream_ptr()), 'actor_grad')
        if return_actions:
            return actor_model._grad_views, actions
        return actor_model._grad_views
