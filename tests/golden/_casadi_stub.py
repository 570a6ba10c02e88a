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


class E:
    """Scalar expression node."""
    __array_priority__ = 1000

    def __init__(self, op, *args):
        self.op, self.args = op, args

    @staticmethod
    def wrap(x):
        return x if isinstance(x, E) else E('const', float(x))

    def ev(self, env):
        if self.op == 'const':
            return self.args[0]
        if self.op == 'sym':
            return env.get(self.args[0], 0.0)            # unbound symbols: the cost is separable in x and u
        if self.op == 'call':                            # Function applied to symbolic arguments
            fn, k, argv = self.args
            inner = {}
            for sym_vec, arg in zip(fn.inputs, argv):
                for s, a in zip(sym_vec.items, arg.items):
                    inner[s.args[0]] = a.ev(env)
            return fn.outputs[0].items[k].ev(inner)
        return _num(self.op, *[a.ev(env) for a in self.args])

    def __add__(self, o): return E('add', self, E.wrap(o))
    def __radd__(self, o): return E('add', E.wrap(o), self)
    def __sub__(self, o): return E('sub', self, E.wrap(o))
    def __rsub__(self, o): return E('sub', E.wrap(o), self)
    def __mul__(self, o): return E('mul', self, E.wrap(o))
    def __rmul__(self, o): return E('mul', E.wrap(o), self)
    def __truediv__(self, o): return E('div', self, E.wrap(o))
    def __rtruediv__(self, o): return E('div', E.wrap(o), self)
    def __pow__(self, n): return E('pow', self, E.wrap(n))
    def __neg__(self): return E('neg', self)
    def log(self): return E('log', self)
    def exp(self): return E('exp', self)
    def sqrt(self): return E('sqrt', self)
    def cos(self): return E('cos', self)
    def sin(self): return E('sin', self)
    def tan(self): return E('tan', self)


class Vec:
    """Column vector of scalar expressions (casadi.SX n x 1)."""

    def __init__(self, items):
        self.items = list(items)

    def __len__(self): return len(self.items)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return Vec(self.items[k])
        return self.items[k]

    def __setitem__(self, k, v):
        if isinstance(k, slice):
            vals = v.items if isinstance(v, Vec) else list(v)
            self.items[k] = [E.wrap(x) for x in vals]
        else:
            self.items[k] = E.wrap(v)


class Deriv:
    """Lazy derivative of a scalar expression: kind 'grad' (w.r.t. x), 'hess' (x, x) or 'mixed' (x, u)."""

    def __init__(self, kind, expr, x, u=None):
        self.kind, self.expr, self.x, self.u = kind, expr, x, u


class _SX:
    def __call__(self, n, m=1):
        assert m == 1
        return Vec([E('const', 0.0) for _ in range(n)])

    @staticmethod
    def sym(name, n, m=1):
        assert m == 1
        return Vec([E('sym', (name, i, id(object()))) for i in range(n)])


class Function:
    def __init__(self, name, inputs, outputs, *names):
        self.name, self.inputs = name, inputs
        self.outputs = [o if isinstance(o, (Vec, Deriv)) else Vec([o]) for o in outputs]

    def __call__(self, *args):
        symbolic = any(isinstance(a, Vec) or isinstance(a, E) for a in args)
        out = self.outputs[0]
        if symbolic:
            argv = [a if isinstance(a, Vec) else Vec([a]) for a in args]
            assert isinstance(out, Vec)
            res = Vec([E('call', self, k, argv) for k in range(len(out))])
            return res if len(res) > 1 else res.items[0]
        vals = [np.asarray(a, dtype=float).reshape(-1) for a in args]
        env = {}
        for sv, v in zip(self.inputs, vals):
            for s, x in zip(sv.items, v):
                env[s.args[0]] = float(x)
        if isinstance(out, Vec):
            return np.array([[o.ev(env)] for o in out.items], dtype=float)
        xs = out.x.items
        n = len(xs)
        if out.kind in ('grad', 'hess'):
            g, H = np.zeros(n), np.zeros((n, n))
            for i in range(n):
                for j in range(i, n):
                    e2 = dict(env)
                    for k, s in enumerate(xs):
                        e2[s.args[0]] = HD(env.get(s.args[0], 0.0), 1.0 if k == i else 0.0, 1.0 if k == j else 0.0)
                    r = _lift(out.expr.ev(e2))
                    H[i, j] = H[j, i] = r.ab
                    if i == j:
                        g[i] = r.a
            return g.reshape(n, 1) if out.kind == 'grad' else H
        us = out.u.items
        M = np.zeros((n, len(us)))
        for i in range(n):
            for j in range(len(us)):
                e2 = dict(env)
                e2[xs[i].args[0]] = HD(env.get(xs[i].args[0], 0.0), 1.0, 0.0)
                e2[us[j].args[0]] = HD(env.get(us[j].args[0], 0.0), 0.0, 1.0)
                M[i, j] = _lift(out.expr.ev(e2)).ab
        return M


def hessian(expr, x):
    return Deriv('hess', expr, x), Deriv('grad', expr, x)


def jacobian(expr, x):
    if isinstance(expr, Deriv) and expr.kind == 'grad':
        return Deriv('mixed', expr.expr, expr.x, x)
    return Deriv('grad', expr, x)


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
    sys.modules['casadi'] = m
    return m
