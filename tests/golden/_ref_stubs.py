"""Stub third-party modules so that the UNMODIFIED reference sources under /root/reference can be
imported in the build container (TensorFlow / Pinocchio / CasADi / tf_siren are not installable
here).  Used only by tests/golden/make_golden.py and by container-only cross-checks; nothing
here runs on the GPU box.  The stubs implement, with NumPy, exactly the handful of calls the
reference makes on the code paths we execute."""
import os
import sys
import types

import numpy as np

REF = '/root/reference'


def have_reference():
    return os.path.isdir(REF) and os.path.exists(os.path.join(REF, 'environment.py'))


class _MSE:
    def __init__(self, reduction=None):
        self.reduction = reduction

    def __call__(self, a, b, sample_weight=None):
        d = (np.asarray(a, dtype=np.float32) - np.asarray(b, dtype=np.float32)) ** 2
        return _T(d.mean(axis=-1))


class _T(np.ndarray):
    """ndarray with the .numpy()/.shape surface of an eager tensor."""
    def __new__(cls, a):
        return np.asarray(a).view(cls)

    def numpy(self):
        return np.asarray(self)


def install():
    """Insert stub modules into sys.modules and put the reference on sys.path."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if not hasattr(np, 'math'):          # numpy 1.24 (the reference's pin) still had np.math = math
        import math
        np.math = math
    tf = types.ModuleType('tensorflow')
    tf.float32 = np.float32
    tf.float64 = np.float64
    tf.cos = lambda x: float(np.cos(x))
    tf.sin = lambda x: float(np.sin(x))
    tf.convert_to_tensor = lambda x, dtype=None: _T(np.asarray(x, dtype=dtype))
    tf.reduce_sum = lambda x, axis=None: _T(np.sum(np.asarray(x), axis=axis, dtype=np.asarray(x).dtype))
    tf.reshape = lambda x, shape: _T(np.reshape(np.asarray(x), shape))
    tf.is_tensor = lambda x: isinstance(x, _T)
    tf.function = lambda f: f
    tf.squeeze = lambda x: _T(np.squeeze(np.asarray(x)))
    tf.math = types.SimpleNamespace(abs=lambda x: _T(np.abs(x)), subtract=lambda a, b: _T(np.asarray(a) - np.asarray(b)))
    keras = types.ModuleType('tensorflow.keras')
    keras.losses = types.SimpleNamespace(MeanSquaredError=_MSE, Reduction=types.SimpleNamespace(NONE='none'))
    keras.layers = types.ModuleType('tensorflow.keras.layers')
    keras.regularizers = types.ModuleType('tensorflow.keras.regularizers')
    tf.keras = keras
    sys.modules['tensorflow'] = tf
    sys.modules['tensorflow.keras'] = keras
    sys.modules['tensorflow.keras.layers'] = keras.layers
    sys.modules['tensorflow.keras.regularizers'] = keras.regularizers
    siren = types.ModuleType('tf_siren')
    siren.SinusodialRepresentationDense = object
    sys.modules['tf_siren'] = siren

    pin = types.ModuleType('pinocchio')
    cpin = types.ModuleType('pinocchio.casadi')

    class _M:
        def __init__(self, *a):
            pass

        def createData(self):
            return None
    cpin.Model = _M
    pin.casadi = cpin
    sys.modules['pinocchio'] = pin
    sys.modules['pinocchio.casadi'] = cpin

    nq = {'double_integrator.urdf': 2, 'planar_manipulator_3dof.urdf': 3, 'ur5_robot.urdf': 6}

    class _Robot:
        def __init__(self, path):
            self.nq = self.nv = self.na = nq[os.path.basename(path)]
            self.model = None

    class _RW:
        @staticmethod
        def BuildFromURDF(path, dirs):
            return _Robot(path)

    class _RS:
        def __init__(self, *a, **k):
            pass
    ru = types.ModuleType('robot_utils')
    ru.RobotWrapper = _RW
    ru.RobotSimulator = _RS
    sys.modules['robot_utils'] = ru
    return tf


def import_conf(system_id):
    import importlib
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        return importlib.import_module('conf_' + system_id)
    finally:
        os.chdir(cwd)
