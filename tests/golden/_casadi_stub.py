"""A small symbolic stand-in for the handful of CasADi calls on the reference's backward_pass path (TO.py:119-202 and the
``*_CAMS`` cost / p_ee / simulate functions of environment_TO.py), so that the UNMODIFIED reference code can be executed in the
build container where ``casadi==3.6.3`` is not installable.  Expressions are kept as a graph; ``hessian`` / ``jacobian`` are
lazy and evaluated exactly with hyper-dual numbers when a ``Function`` is called with numbers.  Build-container only."""
import math
import types

import numpy as np


class HD:
    __slots__ = ('v', 'a', 'b', 'ab')

    def __init__(self, v, a=0.0, b=0.0, ab=0.0):
        self.v, self.a, self.b, self.ab = float(v), float(a), float(b), float(ab)

    def un(self, f0, f1, f2):
        return HD(f0, f1 * self.a, f1 * self.b, f1 * self.ab + f2 * self.a * self.b)


def _lift(x):
    return x if isinstance(x, HD) else HD(x)


def _num(op, *v):
    if op in ('add', 'sub', 'mul', 'div', 'pow'):
        x, y = v
        if not isinstance(x, HD) and not isinstance(y, HD):
            return {'add': x + y, 'sub': x - y, 'mul': x * y, 'div': x / y if op == 'div' else 0, 'pow': x ** y if op == 'pow' else 0}[op]
        if op == 'pow':
            n = float(y.v if isinstance(y, HD) else y)
            x = _lift(x)
            return x.un(x.v ** n, n * x.v ** (n - 1), n * (n - 1) * x.v ** (n - 2))
        x, y = _lift(x), _lift(y)
        if op == 'add':
            return HD(x.v + y.v, x.a + y.a, x.b + y.b, x.ab + y.ab)
        if op == 'sub':
            return HD(x.v - y.v, x.a - y.a, x.b - y.b, x.ab - y.ab)
        if op == 'div':
            r = y.un(1.0 / y.v, -1.0 / y.v ** 2, 2.0 / y.v ** 3)
            y = r
        return HD(x.v * y.v, x.v * y.a + x.a * y.v, x.v * y.b + x.b * y.v, x.v * y.ab + x.a * y.b + x.b * y.a + x.ab * y.v)
    x = v[0]
    if not isinstance(x, HD):
        return {'neg': lambda z: -z, 'log': math.log, 'exp': math.exp, 'sqrt': math.sqrt, 'cos': math.cos, 'sin': math.sin, 'tan': math.tan}[op](x)
    if op == 'neg':
        return HD(-x.v, -x.a, -x.b, -x.ab)
    if op == 'log':
        return x.un(math.log(x.v), 1 / x.v, -1 / x.v ** 2)
    if op == 'exp':
        e = math.exp(x.v)
        return x.un(e, e, e)
    if op == 'sqrt':
        s = math.sqrt(x.v)
        return x.un(s, 0.5 / s, -0.25 / (s * x.v))
    if op == 'cos':
        return x.un(math.cos(x.v), -math.sin(x.v), -math.cos(x.v))
    if op == 'sin':
        return x.un(math.sin(x.v), math.cos(x.v), -math.sin(x.v))
    if op == 'tan':
        t = math.tan(x.v)
        return x.un(t, 1 + t * t, 2 * t * (1 + t * t))
    raise NotImplementedError(op)


_UFUNC = {'add': 'add', 'subtract': 'sub', 'multiply': 'mul', 'true_divide': 'div', 'divide': 'div', 'power': 'pow', 'negative': 'neg',
          'log': 'log', 'exp': 'exp', 'sqrt': 'sqrt', 'cos': 'cos', 'sin': 'sin', 'tan': 'tan'}


def _apply(op, *a):
    a = [float(x) if isinstance(x, (np.floating, np.integer)) or (isinstance(x, np.ndarray) and x.ndim == 0 and x.dtype != object) else
         (x.item() if isinstance(x, np.ndarray) and x.ndim == 0 else x) for x in a]
    if not any(isinstance(x, E) for x in a):
        return _num(op, *a)
    return E(op, *[E.wrap(x) for x in a])


class E:
    """Scalar expression node.  Containers (casadi.SX n x 1, matrices) are plain NumPy object arrays of E / floats."""

    def __init__(self, op, *args):
        self.op, self.args = op, args

    @staticmethod
    def wrap(x):
        return x if isinstance(x, E) else E('const', float(x))

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__' or ufunc.__name__ not in _UFUNC:
            return NotImplemented
        op = _UFUNC[ufunc.__name__]
        if any(isinstance(i, np.ndarray) and i.ndim > 0 for i in inputs):
            return np.frompyfunc(lambda *a: _apply(op, *a), len(inputs), 1)(*inputs)
        return _apply(op, *inputs)

    def ev(self, env):
        if self.op == 'const':
            return self.args[0]
        if self.op == 'sym':
            return env.get(self.args[0], 0.0)            # unbound symbols: the cost is separable in x and u
        if self.op == 'call':                            # Function applied to symbolic arguments
            fn, k, argv = self.args
            inner = {}
            for sym_vec, arg in zip(fn.inputs, argv):
                for s_, a in zip(sym_vec, arg):
                    inner[s_.args[0]] = _ev(a, env)
            return _ev(fn.outputs[0][k], inner)
        return _num(self.op, *[a.ev(env) for a in self.args])

    def _b(self, op, o, swap=False):
        if isinstance(o, np.ndarray) and o.ndim > 0:
            return np.frompyfunc((lambda x: _apply(op, x, self)) if swap else (lambda x: _apply(op, self, x)), 1, 1)(o)
        return _apply(op, o, self) if swap else _apply(op, self, o)

    def __add__(self, o): return self._b('add', o)
    def __radd__(self, o): return self._b('add', o, True)
    def __sub__(self, o): return self._b('sub', o)
    def __rsub__(self, o): return self._b('sub', o, True)
    def __mul__(self, o): return self._b('mul', o)
    def __rmul__(self, o): return self._b('mul', o, True)
    def __truediv__(self, o): return self._b('div', o)
    def __rtruediv__(self, o): return self._b('div', o, True)
    def __pow__(self, n): return self._b('pow', n)
    def __neg__(self): return E('neg', self)
    def log(self): return E('log', self)
    def exp(self): return E('exp', self)
    def sqrt(self): return E('sqrt', self)
    def cos(self): return E('cos', self)
    def sin(self): return E('sin', self)
    def tan(self): return E('tan', self)


def _ev(x, env):
    return x.ev(env) if isinstance(x, E) else float(x)


def _vec(x):
    """1-D object array view of a symbolic vector / scalar / list."""
    if isinstance(x, E):
        a = np.empty(1, dtype=object)
        a[0] = x
        return a
    return np.asarray(x, dtype=object).reshape(-1)


class Deriv:
    """Lazy derivative of a scalar expression: kind 'grad' (w.r.t. x), 'hess' (x, x) or 'mixed' (x, u)."""

    def __init__(self, kind, expr, x, u=None):
        self.kind, self.expr, self.x, self.u = kind, expr, x, u


class _SX:
    def __call__(self, n, m=1):
        assert m == 1
        a = np.empty(n, dtype=object)
        a[:] = 0.0
        return a

    @staticmethod
    def sym(name, n, m=1):
        assert m == 1
        a = np.empty(n, dtype=object)
        for i in range(n):
            a[i] = E('sym', (name, i, id(a)))
        return a


class Function:
    def __init__(self, name, inputs, outputs, *names):
        self.name, self.inputs = name, [_vec(i) for i in inputs]
        self.outputs = [o if isinstance(o, Deriv) else _vec(o) for o in outputs]

    def __call__(self, *args):
        symbolic = any(isinstance(a, E) or (isinstance(a, np.ndarray) and a.dtype == object) for a in args)
        out = self.outputs[0]
        if symbolic:
            argv = [_vec(a) for a in args]
            assert not isinstance(out, Deriv)
            res = np.empty(len(out), dtype=object)
            for k in range(len(out)):
                res[k] = E('call', self, k, argv)
            return res if len(res) > 1 else res[0]
        vals = [np.asarray(a, dtype=float).reshape(-1) for a in args]
        env = {}
        for sv, v in zip(self.inputs, vals):
            for s_, x in zip(sv, v):
                env[s_.args[0]] = float(x)
        if not isinstance(out, Deriv):
            return np.array([[_ev(o, env)] for o in out], dtype=float)
        xs = _vec(out.x)
        n = len(xs)
        if out.kind in ('grad', 'hess'):
            g, H = np.zeros(n), np.zeros((n, n))
            for i in range(n):
                for j in range(i, n):
                    e2 = dict(env)
                    for k, s_ in enumerate(xs):
                        e2[s_.args[0]] = HD(env.get(s_.args[0], 0.0), 1.0 if k == i else 0.0, 1.0 if k == j else 0.0)
                    r = _lift(_ev(out.expr, e2))
                    H[i, j] = H[j, i] = r.ab
                    if i == j:
                        g[i] = r.a
            return g.reshape(n, 1) if out.kind == 'grad' else H
        us = _vec(out.u)
        M = np.zeros((n, len(us)))
        for i in range(n):
            for j in range(len(us)):
                e2 = dict(env)
                e2[xs[i].args[0]] = HD(env.get(xs[i].args[0], 0.0), 1.0, 0.0)
                e2[us[j].args[0]] = HD(env.get(us[j].args[0], 0.0), 0.0, 1.0)
                M[i, j] = _lift(_ev(out.expr, e2)).ab
        return M


def _scalar(expr):
    if isinstance(expr, np.ndarray):
        assert expr.size == 1
        return expr.reshape(-1)[0]
    return expr


def hessian(expr, x):
    return Deriv('hess', _scalar(expr), x), Deriv('grad', _scalar(expr), x)


def jacobian(expr, x):
    if isinstance(expr, Deriv) and expr.kind == 'grad':
        return Deriv('mixed', expr.expr, expr.x, x)
    return Deriv('grad', _scalar(expr), x)


def install():
    import sys
    m = types.ModuleType('casadi')
    m.SX = _SX()
    m.Function = Function
    m.hessian = hessian
    m.jacobian = jacobian
    m.cos = lambda e: E.wrap(e).cos()
    m.sin = lambda e: E.wrap(e).sin()
    m.tan = lambda e: E.wrap(e).tan()
    m.horzcat = lambda *a: np.array(list(a), dtype=object).reshape(1, -1)
    m.vertcat = lambda *a: np.vstack([np.atleast_2d(np.asarray(x, dtype=object)) for x in a])
    m.mtimes = lambda a, b: np.dot(np.asarray(a, dtype=object), np.asarray(b, dtype=object))
    m.repmat = lambda a, r, c: np.tile(np.atleast_2d(np.asarray(a, dtype=object)), (r, c))
    m.sum1 = lambda a: np.sum(np.asarray(a, dtype=object), axis=0)
    sys.modules['casadi'] = m
    return m
