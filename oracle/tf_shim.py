"""TEST INFRASTRUCTURE ONLY -- a minimal eager stand-in for the `tensorflow` module, just large enough to execute
the VERBATIM agent code of the reference (objects.py: Network / Critic / Actor / Critic_big / Actor_big / IDHPsp /
IDHPnonlin) in a container where TensorFlow cannot be installed.

What it pins and what it does not.  Running the reference's own classes on this stand-in pins the reference's CONTROL
FLOW, CALL ORDER, ALIASING and every numpy-side expression (traces, RLS, adaptation logic, logging) -- they execute
unmodified, on real numpy.  The arithmetic INSIDE TensorFlow ops is not observable here; the stand-in implements the
contract stated in DESIGN.md section 3 (the same one the C oracle and the CUDA kernels follow):

  * tensors are float32; a numpy / python operand meeting a tensor is converted to float32 first (TensorFlow's
    `convert_to_tensor(..., dtype_hint=float32)`); `__array_priority__ = 100` makes numpy defer to the tensor's
    reflected operators as TensorFlow's EagerTensor does;
  * every element-wise op rounds once; `@` is an in-order FMA chain over the contraction index;
  * reverse-mode autodiff uses TensorFlow's gradient definitions: TanhGrad = dy * (1 - y*y),
    MatMul grads dA = dC @ B^T, dB = A^T @ dC (same chain order), Reshape / Transpose / Split / StridedSlice = data
    movement;
  * SGD: var <- var - lr * grad (two roundings, float32 learning-rate variable);
  * tanh is pluggable (`set_tanh`): libm tanhf or the t13 kernel the CUDA path uses;
  * `tf.random.normal` returns `noise * stddev + mean` with `noise` popped from an injected stream (`set_noise`).

Never imported by the product package.
"""
from __future__ import annotations

import ctypes
import sys
import types

import numpy as np

_libm = ctypes.CDLL("libm.so.6")
_libm.fmaf.restype = ctypes.c_float
_libm.fmaf.argtypes = [ctypes.c_float] * 3
_libm.tanhf.restype = ctypes.c_float
_libm.tanhf.argtypes = [ctypes.c_float]

F32 = np.float32
float32 = "float32"          # tf.float32 token

_state = {"tanh": None, "noise": None, "noise_pos": 0, "tapes": []}


def set_tanh(fn):
    """fn: float32 ndarray -> float32 ndarray (element-wise)."""
    _state["tanh"] = fn


def set_noise(stream):
    """stream: 1-D float32 array of N(0,1) draws, consumed one per tf.random.normal call."""
    _state["noise"] = None if stream is None else np.asarray(stream, dtype=F32)
    _state["noise_pos"] = 0


def _tanh(x):
    if _state["tanh"] is not None:
        return np.asarray(_state["tanh"](np.asarray(x, dtype=F32)), dtype=F32)
    flat = np.asarray(x, dtype=F32).ravel()
    return np.array([_libm.tanhf(float(v)) for v in flat], dtype=F32).reshape(np.shape(x))


def _chain_matmul(A, B):
    """(m,k) @ (k,n), float32, in-order FMA chain over k (first term a plain product)."""
    A = np.asarray(A, dtype=F32); B = np.asarray(B, dtype=F32)
    assert A.ndim == 2 and B.ndim == 2 and A.shape[1] == B.shape[0], (A.shape, B.shape)
    out = np.empty((A.shape[0], B.shape[1]), dtype=F32)
    for i in range(A.shape[0]):
        for j in range(B.shape[1]):
            acc = F32(A[i, 0] * B[0, j])
            for t in range(1, A.shape[1]):
                acc = F32(_libm.fmaf(float(A[i, t]), float(B[t, j]), float(acc)))
            out[i, j] = acc
    return out


def _unbroadcast(g, shape):
    g = np.asarray(g, dtype=F32)
    while g.ndim > len(shape):
        g = g.sum(axis=0, dtype=F32)
    for ax, n in enumerate(shape):
        if n == 1 and g.shape[ax] != 1:
            g = g.sum(axis=ax, keepdims=True, dtype=F32)
    return g.reshape(shape)


class Tensor:
    __array_priority__ = 100

    def __init__(self, value, parents=()):
        self.v = np.array(value, dtype=F32)
        self._parents = tuple(parents) if _state["tapes"] else ()     # (parent Tensor, vjp) pairs, recorded under a tape

    # ---- conversions
    def numpy(self):
        return self.v.copy() if self.v.ndim else F32(self.v)

    def __array__(self, dtype=None, copy=None):
        return self.v.astype(dtype) if dtype is not None else self.v.copy()

    @property
    def shape(self):
        return tuple(self.v.shape)

    @property
    def dtype(self):
        return float32

    def __len__(self):
        return self.v.shape[0]

    def __float__(self):
        return float(self.v)

    def __repr__(self):
        return f"<shim.Tensor {self.v!r}>"

    # ---- element-wise
    def _bin(self, other, fn, vjp_a, vjp_b, swap=False):
        o = _as_tensor(other)
        a, b = (o, self) if swap else (self, o)
        out = fn(a.v, b.v).astype(F32)
        return Tensor(out, [(a, lambda g: _unbroadcast(vjp_a(np.asarray(g, dtype=F32), a.v, b.v, out), a.v.shape)),
                            (b, lambda g: _unbroadcast(vjp_b(np.asarray(g, dtype=F32), a.v, b.v, out), b.v.shape))])

    def __add__(self, o): return self._bin(o, lambda a, b: a + b, lambda g, a, b, y: g, lambda g, a, b, y: g)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, lambda g, a, b, y: g, lambda g, a, b, y: g, swap=True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b, lambda g, a, b, y: g, lambda g, a, b, y: -g)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, lambda g, a, b, y: g, lambda g, a, b, y: -g, swap=True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b, lambda g, a, b, y: g * b, lambda g, a, b, y: g * a)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, lambda g, a, b, y: g * b, lambda g, a, b, y: g * a, swap=True)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b, lambda g, a, b, y: g / b, lambda g, a, b, y: -(g * a) / (b * b))

    def __neg__(self):
        return Tensor(-self.v, [(self, lambda g: -np.asarray(g, dtype=F32))])

    def __pow__(self, p):
        assert p == 2, "only squaring is used by objects.py"
        return Tensor(self.v * self.v, [(self, lambda g: (np.asarray(g, dtype=F32) * F32(2)) * self.v)])

    # ---- matmul
    def __matmul__(self, o):
        return matmul(self, o)

    def __rmatmul__(self, o):
        return matmul(o, self)

    # ---- indexing
    def __getitem__(self, idx):
        shape = self.v.shape

        def vjp(g):
            z = np.zeros(shape, dtype=F32)
            z[idx] = g
            return z
        return Tensor(self.v[idx], [(self, vjp)])


class Variable(Tensor):
    """tf.Variable: same object for the life of the layer; assign() mutates in place."""

    def __init__(self, value):
        super().__init__(value)
        self._parents = ()

    def assign(self, value):
        self.v = np.array(_as_tensor(value).v, dtype=F32).reshape(self.v.shape)
        return self

    def assign_sub(self, value):
        self.v = (self.v - _as_tensor(value).v).astype(F32)
        return self


def _as_tensor(x):
    """TensorFlow's convert_to_tensor(x, dtype_hint=float32): tensors pass through, float32 arrays are taken as they are,
    everything else (python floats, float64 arrays / scalars) is rounded to float32 once."""
    if isinstance(x, Tensor):
        return x
    if isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in x):
        return Tensor(np.array([np.asarray(e.v if isinstance(e, Tensor) else e) for e in x], dtype=F32))
    if isinstance(x, np.ndarray) and x.dtype == F32:
        return Tensor(x)
    return Tensor(np.asarray(x, dtype=np.float64).astype(F32))


def matmul(a, b):
    a, b = _as_tensor(a), _as_tensor(b)
    return Tensor(_chain_matmul(a.v, b.v), [(a, lambda g: _chain_matmul(g, b.v.T)), (b, lambda g: _chain_matmul(a.v.T, g))])


def constant(value, dtype=None):
    return Tensor(np.asarray(value, dtype=np.float64).astype(F32) if not isinstance(value, Tensor) else value.v)


def convert_to_tensor(value, dtype=None, dtype_hint=None):
    return value if isinstance(value, Tensor) else constant(value)


def reshape(x, shape):
    if isinstance(x, (list, tuple)):                     # tf.reshape([tensor, ...], shape) packs the list first
        parts = [_as_tensor(e) for e in x]
        src_shapes = [p.v.shape for p in parts]
        packed = np.stack([p.v for p in parts])

        def mk(i):
            return lambda g: np.asarray(g, dtype=F32).reshape(packed.shape)[i].reshape(src_shapes[i])
        return Tensor(packed.reshape(shape), [(p, mk(i)) for i, p in enumerate(parts)])
    x = _as_tensor(x)
    src = x.v.shape
    return Tensor(x.v.reshape(shape), [(x, lambda g: np.asarray(g, dtype=F32).reshape(src))])


def transpose(x):
    x = _as_tensor(x)
    return Tensor(x.v.T, [(x, lambda g: np.asarray(g, dtype=F32).T)])


def split(x, n, axis=0):
    x = _as_tensor(x)
    outs = []
    for i, piece in enumerate(np.split(x.v, n, axis=axis)):
        sl = [slice(None)] * x.v.ndim
        w = x.v.shape[axis] // n
        sl[axis] = slice(i * w, (i + 1) * w)

        def vjp(g, sl=tuple(sl)):
            z = np.zeros(x.v.shape, dtype=F32)
            z[sl] = g
            return z
        outs.append(Tensor(piece, [(x, vjp)]))
    return outs


def sqrt(x):
    x = _as_tensor(x)
    y = np.sqrt(x.v).astype(F32)
    return Tensor(y, [(x, lambda g: (np.asarray(g, dtype=F32) * F32(0.5)) / y)])


def tanh(x):
    x = _as_tensor(x)
    y = _tanh(x.v)
    return Tensor(y, [(x, lambda g: np.asarray(g, dtype=F32) * (F32(1) - y * y))])      # TanhGrad


def _linear(x):
    return x


class GradientTape:
    def __init__(self, persistent=False):
        self.persistent = persistent

    def __enter__(self):
        _state["tapes"].append(self)
        return self

    def __exit__(self, *exc):
        _state["tapes"].remove(self)
        return False

    def watch(self, t):
        pass                                               # every op under an active tape is recorded

    def gradient(self, target, sources, output_gradients=None):
        targets = target if isinstance(target, (list, tuple)) else [target]
        single = not isinstance(sources, (list, tuple))
        srcs = [sources] if single else list(sources)
        grads = {}
        order, seen = [], set()

        def visit(t):
            if id(t) in seen:
                return
            seen.add(id(t))
            for p, _ in t._parents:
                visit(p)
            order.append(t)
        for t in targets:
            visit(t)
        for i, t in enumerate(targets):
            og = output_gradients[i] if isinstance(output_gradients, (list, tuple)) else output_gradients
            if og is None:
                seed = np.ones(t.v.shape, dtype=F32)
            else:
                seed = np.asarray(og.v if isinstance(og, Tensor) else og, dtype=np.float64).astype(F32).reshape(t.v.shape)
            grads[id(t)] = seed if id(t) not in grads else (grads[id(t)] + seed).astype(F32)
        for t in reversed(order):
            g = grads.get(id(t))
            if g is None:
                continue
            for p, vjp in t._parents:
                gp = np.asarray(vjp(g), dtype=F32)
                grads[id(p)] = gp if id(p) not in grads else (grads[id(p)] + gp).astype(F32)
        out = [None if id(s) not in grads else Tensor(grads[id(s)]) for s in srcs]
        return out[0] if single else out


# ------------------------------------------------------------------ keras
class _Dense:
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None):
        assert not use_bias
        self.units = units
        self.activation = tanh if activation == "tanh" else _linear
        self.kernel_initializer = kernel_initializer
        self.weights = []

    def build(self, in_dim):
        self.kernel = Variable(self.kernel_initializer((in_dim, self.units)))
        self.weights = [self.kernel]


class _Sequential:
    def __init__(self):
        self.layers = []
        self._in_dim = None

    def add(self, layer):
        if isinstance(layer, _InputSpec):
            self._in_dim = layer.dim
            return
        layer.build(self._in_dim)
        self._in_dim = layer.units
        self.layers.append(layer)


class _InputSpec:
    def __init__(self, shape):
        self.dim = shape[0]


class _Model:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self.call(*a, **k)

    @property
    def trainable_weights(self):
        out = []
        for v in self.__dict__.values():
            if isinstance(v, _Sequential):
                for layer in v.layers:
                    out.extend(layer.weights)
        return out

    def get_weights(self):
        return [w.numpy() for w in self.trainable_weights]

    def set_weights(self, weights):
        for var, w in zip(self.trainable_weights, weights):
            var.assign(np.asarray(w, dtype=F32))


class _TruncatedNormal:
    def __init__(self, stddev=0.05, seed=None, mean=0.0):
        self.stddev, self.rng = stddev, np.random.RandomState(seed)

    def __call__(self, shape):
        out = self.rng.standard_normal(shape)
        bad = np.abs(out) > 2.0
        while bad.any():
            out[bad] = self.rng.standard_normal(int(bad.sum()))
            bad = np.abs(out) > 2.0
        return (out * self.stddev).astype(F32)


class _Identity:
    def __call__(self, shape):
        return np.eye(*shape, dtype=F32)


class _SGD:
    def __init__(self, learning_rate=0.01):
        self.learning_rate = Variable(np.asarray(learning_rate, dtype=np.float64).astype(F32))

    def apply_gradients(self, grads_and_vars):
        lr = self.learning_rate.v
        for g, var in grads_and_vars:
            gv = g.v if isinstance(g, Tensor) else np.asarray(g, dtype=F32)
            var.assign((var.v - (lr * gv).astype(F32)).astype(F32))


def _random_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None):
    assert _state["noise"] is not None, "tf_shim.set_noise(stream) first"
    z = np.full(shape, _state["noise"][_state["noise_pos"]], dtype=F32)
    _state["noise_pos"] += 1
    sd = np.asarray(stddev, dtype=np.float64).astype(F32)
    m = mean.v if isinstance(mean, Tensor) else np.asarray(mean, dtype=np.float64).astype(F32)
    return Tensor(((z * sd).astype(F32) + m).astype(F32))


def install():
    """Register this module as `tensorflow` (idempotent) and return it."""
    me = sys.modules[__name__]
    keras = types.SimpleNamespace(
        Model=_Model, Sequential=_Sequential, Input=lambda shape: _InputSpec(shape),
        layers=types.SimpleNamespace(Dense=_Dense),
        initializers=types.SimpleNamespace(Identity=_Identity, truncated_normal=_TruncatedNormal, TruncatedNormal=_TruncatedNormal),
        optimizers=types.SimpleNamespace(SGD=_SGD),
        utils=types.SimpleNamespace(set_random_seed=lambda seed: None))
    me.keras = keras
    me.config = types.SimpleNamespace(set_visible_devices=lambda *a, **k: None)
    me.random = types.SimpleNamespace(set_seed=lambda seed: None, normal=_random_normal)
    me.math = types.SimpleNamespace(tanh=tanh, sqrt=sqrt)
    sys.modules["tensorflow"] = me
    return me
