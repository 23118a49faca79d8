"""A minimal stand-in for the `tensorflow` module, backed by torch float64 on the CPU.

TEST INFRASTRUCTURE ONLY.  TensorFlow is not installable in this image (no wheel, no network), so the
reference (`/root/reference/tfrt`) cannot be imported as is.  Its hot path, however, only uses ~70
elementwise / gather / mask / reduce ops of the `tf` namespace, all IEEE-754 float64 on the CPU.  This
module maps each of them onto the torch op with the same documented semantics, so that the reference's
OWN, UNMODIFIED Python source (`tfrt/geometry.py`, `engine.py`, `operation.py`, `materials.py`,
`boundaries.py`, `sources.py`, ...) can be executed in the build container to generate golden vectors
(`tests/golden/make_reference_golden.py`).  `tf.GradientTape` is served by torch autograd.

Nothing in the product or in the GPU tests imports this; the generated `.npz` fixtures travel instead.
Semantics that matter for parity and where they come from (TensorFlow API docs):
  * `tf.math.mod` / `floormod`: result has the sign of the divisor        -> torch.remainder
  * `tf.math.l2_normalize(x, axis)`: x * rsqrt(max(sum(x^2), 1e-12))
  * `tf.linalg.normalize(x, axis=a)`: (x / ||x||, ||x||)
  * `tf.argmin`: lowest index on ties (Eigen CPU kernel)                   -> torch.argmin (first minimal)
  * `tf.meshgrid(a, b)` default indexing "xy": outputs of shape (len(b), len(a))
  * `tf.where(c, x, y)`: select; gradient flows only to the selected branch (same in torch)
  * `tape.gradient(non_scalar, vars)`: gradient of the sum
"""
import builtins
import contextlib
import math as _math

import numpy as _np
import torch as _torch

__version__ = "2.shim"

float64 = _torch.float64
float32 = _torch.float32
int64 = _torch.int64
int32 = _torch.int32
bool = _torch.bool  # noqa: A001  (mirrors tf.bool)
string = str
DType = _torch.dtype
dtype = _torch.dtype
Tensor = _torch.Tensor


def _t(x, dtype=None):
    if isinstance(x, _torch.Tensor):
        return x if dtype is None or x.dtype == dtype else x.to(dtype)
    if isinstance(x, _np.ndarray):
        out = _torch.from_numpy(_np.ascontiguousarray(x))
        return out if dtype is None else out.to(dtype)
    if isinstance(x, (list, tuple)) and len(x) > 0 and any(isinstance(e, _torch.Tensor) for e in x):
        return _torch.stack([_t(e, dtype) for e in x])
    if dtype is None:
        if isinstance(x, builtins.bool):
            dtype = _torch.bool
        elif isinstance(x, int):
            dtype = _torch.int32
        elif isinstance(x, float):
            dtype = _torch.float32
        else:
            arr = _np.asarray(x)
            out = _torch.from_numpy(_np.ascontiguousarray(arr))
            if arr.dtype == _np.float64 and not isinstance(x, _np.ndarray):
                # python floats in lists default to float32 in TF; the reference always casts or
                # passes dtype, so keep float64 (harmless: wider)
                return out
            return out
    return _torch.as_tensor(x, dtype=dtype)


def _pair(a, b):
    """Binary-op operand promotion the way TF does it for python scalars (take the tensor's dtype)."""
    ta, tb = isinstance(a, _torch.Tensor), isinstance(b, _torch.Tensor)
    if ta and not tb:
        return a, (_t(b, a.dtype) if not isinstance(b, _np.ndarray) else _t(b))
    if tb and not ta:
        return (_t(a, b.dtype) if not isinstance(a, _np.ndarray) else _t(a)), b
    if not ta and not tb:
        a = _t(a)
        return a, _t(b, a.dtype)
    return a, b


def _shape_arg(shape):
    if isinstance(shape, _torch.Tensor):
        return tuple(int(v) for v in shape.reshape(-1).tolist())
    if isinstance(shape, (int, _np.integer)):
        return (int(shape),)
    return tuple(int(v) for v in shape)


class TensorShape(tuple):
    pass


class TensorSpec:
    def __init__(self, shape=None, dtype=None, name=None):
        self.shape, self.dtype, self.name = shape, dtype, name


def function(func=None, **kwargs):
    if func is not None and callable(func):
        return func
    return lambda f: f


@contextlib.contextmanager
def name_scope(name):
    yield


class _Errors:
    class InvalidArgumentError(Exception):
        pass


errors = _Errors()


# ---------------------------------------------------------------------------------- creation
def constant(value, dtype=None, shape=None, name=None):
    out = _t(value, dtype)
    if shape is not None:
        out = out.broadcast_to(_shape_arg(shape)).clone()
    return out


def convert_to_tensor(value, dtype=None):
    return _t(value, dtype)


def cast(x, dtype, name=None):
    return _t(x).to(dtype)


def zeros(shape, dtype=float32):
    return _torch.zeros(_shape_arg(shape), dtype=dtype)


def ones(shape, dtype=float32):
    return _torch.ones(_shape_arg(shape), dtype=dtype)


def fill(dims, value):
    v = _t(value)
    return _torch.full(_shape_arg(dims), v.item(), dtype=v.dtype)


def zeros_like(x, dtype=None, name=None):
    return _torch.zeros_like(_t(x), dtype=dtype)


def ones_like(x, dtype=None, name=None):
    return _torch.ones_like(_t(x), dtype=dtype)


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001
    def _s(v):
        return v.item() if isinstance(v, _torch.Tensor) else v
    if limit is None:
        start, limit = 0, start
    start, limit, delta = _s(start), _s(limit), _s(delta)
    if dtype is None:
        dtype = _torch.int32 if all(isinstance(v, (int, _np.integer)) for v in (start, limit, delta)) \
            else _torch.float32
    return _torch.arange(start, limit, delta, dtype=dtype)


def linspace(start, stop, num):
    s, e = _t(start), _t(stop)
    dt = s.dtype if s.dtype.is_floating_point else (e.dtype if e.dtype.is_floating_point else float32)
    return _torch.linspace(float(s), float(e), int(num), dtype=dt)


def shape(x, out_type=int32):
    return _torch.tensor(list(_t(x).shape), dtype=out_type)


def rank(x):
    return _t(x).dim()


# ---------------------------------------------------------------------------------- elementwise
def where(condition, x=None, y=None, name=None):
    c = _t(condition)
    if x is None:
        return _torch.nonzero(c)
    x, y = _pair(x, y)
    return _torch.where(c, x, y)


def _cmp(fn):
    def op(a, b, name=None):
        a, b = _pair(a, b)
        return fn(a, b)
    return op


less = _cmp(_torch.lt)
less_equal = _cmp(_torch.le)
greater = _cmp(_torch.gt)
greater_equal = _cmp(_torch.ge)
equal = _cmp(_torch.eq)
not_equal = _cmp(_torch.ne)
logical_and = _cmp(_torch.logical_and)
logical_or = _cmp(_torch.logical_or)
atan2 = _cmp(_torch.atan2)


def logical_not(x, name=None):
    return _torch.logical_not(_t(x))


def _un(fn):
    return lambda x, name=None: fn(_t(x))


abs = _un(_torch.abs)  # noqa: A001
sqrt = _un(_torch.sqrt)
sin = _un(_torch.sin)
cos = _un(_torch.cos)
asin = _un(_torch.asin)
acos = _un(_torch.acos)
sign = _un(_torch.sign)


def clip_by_value(x, lo, hi):
    x = _t(x)
    return _torch.minimum(_torch.maximum(x, _t(lo, x.dtype)), _t(hi, x.dtype))


def stop_gradient(x):
    return _t(x).detach()


class _Math:
    @staticmethod
    def mod(a, b, name=None):
        a, b = _pair(a, b)
        return _torch.remainder(a, b)

    floormod = mod

    @staticmethod
    def l2_normalize(x, axis=None, epsilon=1e-12):
        x = _t(x)
        sq = (x * x).sum(dim=axis, keepdim=True) if axis is not None else (x * x).sum()
        return x * _torch.rsqrt(_torch.clamp(sq, min=epsilon))

    @staticmethod
    def is_finite(x):
        return _torch.isfinite(_t(x))

    @staticmethod
    def squared_difference(a, b):
        a, b = _pair(a, b)
        return (a - b) * (a - b)

    sqrt = staticmethod(sqrt)
    abs = staticmethod(abs)
    atan2 = staticmethod(atan2)


math = _Math()


class _Linalg:
    @staticmethod
    def normalize(x, ord="euclidean", axis=None):
        x = _t(x)
        n = _torch.sqrt((x * x).sum(dim=axis, keepdim=True))
        return x / n, n

    @staticmethod
    def cross(a, b):
        return _torch.linalg.cross(_t(a), _t(b), dim=-1)


linalg = _Linalg()


def norm(x, axis=None):
    x = _t(x)
    return _torch.sqrt((x * x).sum(dim=axis)) if axis is not None else _torch.sqrt((x * x).sum())


# ---------------------------------------------------------------------------------- reductions
def _red(fn, empty=None):
    def op(x, axis=None, keepdims=False, name=None):
        x = _t(x)
        if axis is None:
            return fn(x)
        out = fn(x, dim=axis, keepdim=keepdims)
        return out[0] if isinstance(out, tuple) else out
    return op


def reduce_max(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return x.max() if axis is None else x.max(dim=axis, keepdim=keepdims)[0]


def reduce_min(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return x.min() if axis is None else x.min(dim=axis, keepdim=keepdims)[0]


reduce_sum = _red(_torch.sum)
reduce_mean = _red(_torch.mean)
reduce_prod = _red(_torch.prod)
reduce_any = _red(_torch.any)
reduce_all = _red(_torch.all)


def argmin(x, axis=0, output_type=int64, name=None):
    return _torch.argmin(_t(x), dim=axis).to(output_type)


# ---------------------------------------------------------------------------------- shape / index
def gather(params, indices, axis=0, **kwargs):
    params, indices = _t(params), _t(indices).long()
    out = _torch.index_select(params, axis, indices.reshape(-1))
    shp = list(params.shape)
    return out.reshape(shp[:axis] + list(indices.shape) + shp[axis + 1:])


def gather_nd(params, indices, name=None):
    params, indices = _t(params), _t(indices).long()
    return params[tuple(indices[..., i] for i in builtins.range(indices.shape[-1]))]


def boolean_mask(tensor, mask, axis=None, name=None):
    return _t(tensor)[_t(mask).to(_torch.bool)]


def stack(values, axis=0, name=None):
    vals = [_t(v) for v in values]
    dt = vals[0].dtype
    for v in vals:
        dt = _torch.promote_types(dt, v.dtype)
    return _torch.stack([v.to(dt) for v in vals], dim=axis)


def unstack(value, axis=0, num=None):
    return list(_torch.unbind(_t(value), dim=axis))


def concat(values, axis=0, name=None):
    return _torch.cat([_t(v) for v in values], dim=axis)


def reshape(x, shape, name=None):
    return _t(x).reshape(_shape_arg(shape))


def transpose(x, perm=None):
    x = _t(x)
    return x.permute(*perm) if perm is not None else x.permute(*reversed(builtins.range(x.dim())))


def broadcast_to(x, shape, name=None):
    return _t(x).broadcast_to(_shape_arg(shape))


def meshgrid(*args, indexing="xy"):
    if not args:
        return []
    return list(_torch.meshgrid(*[_t(a) for a in args], indexing=indexing))


def repeat(x, repeats, axis=None):
    x = _t(x)
    if isinstance(repeats, (int, _np.integer)):
        return _torch.repeat_interleave(x, int(repeats), dim=axis)
    return _torch.repeat_interleave(x, _t(repeats).long(), dim=axis)


def pad(x, paddings, mode="CONSTANT", constant_values=0):
    x = _t(x)
    p = _t(paddings).long().tolist()
    flat = []
    for lo, hi in reversed(p):
        flat += [lo, hi]
    return _torch.nn.functional.pad(x, flat, value=constant_values)


def matmul(a, b):
    return _torch.matmul(_t(a), _t(b))


# ---------------------------------------------------------------------------------- random / debug
class _Random:
    @staticmethod
    def uniform(shape, minval=0, maxval=1, dtype=float32, seed=None):
        lo = float(minval)
        hi = float(maxval)
        return _torch.rand(_shape_arg(shape), dtype=dtype) * (hi - lo) + lo

    @staticmethod
    def normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None):
        return _torch.randn(_shape_arg(shape), dtype=dtype) * float(stddev) + float(mean)

    @staticmethod
    def set_seed(seed):
        _torch.manual_seed(seed)


random = _Random()
random_uniform = _Random.uniform


class _Debugging:
    @staticmethod
    def _check(ok, message):
        if not builtins.bool(_torch.all(_t(ok))):
            raise errors.InvalidArgumentError(message or "assertion failed")

    def assert_greater(self, x, y, message=None):
        self._check(greater(x, y), message)

    def assert_greater_equal(self, x, y, message=None):
        self._check(greater_equal(x, y), message)

    def assert_less(self, x, y, message=None):
        self._check(less(x, y), message)

    def assert_less_equal(self, x, y, message=None):
        self._check(less_equal(x, y), message)

    def assert_positive(self, x, message=None):
        self._check(greater(x, 0), message)

    def assert_integer(self, x, message=None):
        if _t(x).dtype.is_floating_point:
            raise TypeError(message or "expected an integer tensor")


debugging = _Debugging()


# ---------------------------------------------------------------------------------- variables, tape
def Variable(initial_value, dtype=None, validate_shape=True, trainable=True, name=None):
    """A leaf tensor that requires grad; `assign*` are patched onto torch.Tensor below."""
    v = _t(initial_value, dtype).detach().clone()
    if v.dtype.is_floating_point:
        v.requires_grad_(trainable)
    return v


def _assign(self, value):
    with _torch.no_grad():
        self.copy_(_t(value, self.dtype))
    return self


def _assign_add(self, value):
    with _torch.no_grad():
        self.add_(_t(value, self.dtype))
    return self


def _assign_sub(self, value):
    with _torch.no_grad():
        self.sub_(_t(value, self.dtype))
    return self


_torch.Tensor.assign = _assign
_torch.Tensor.assign_add = _assign_add
_torch.Tensor.assign_sub = _assign_sub
_orig_numpy = _torch.Tensor.numpy


def _numpy(self, *a, **k):
    return _orig_numpy(self.detach(), *a, **k)


_torch.Tensor.numpy = _numpy


def _reflected(name):
    """TF converts python sequences on the left of a tensor operator (`(0, 0, 4) - tensor`); torch's
    reflected operators return NotImplemented for them."""
    orig = getattr(_torch.Tensor, name)

    def op(self, other):
        if isinstance(other, (list, tuple)):
            other = _torch.as_tensor(other, dtype=self.dtype)
        return orig(self, other)
    return op


for _name in ("__rsub__", "__radd__", "__rmul__", "__rtruediv__"):
    setattr(_torch.Tensor, _name, _reflected(_name))


class GradientTape:
    def __init__(self, persistent=False, watch_accessed_variables=True):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, x):
        pass

    def gradient(self, target, sources, output_gradients=None):
        single = isinstance(sources, _torch.Tensor)
        srcs = [sources] if single else list(sources)
        total = target.sum() if output_gradients is None else (target * output_gradients).sum()
        grads = _torch.autograd.grad(total, srcs, allow_unused=True, retain_graph=self.persistent)
        grads = list(grads)
        return grads[0] if single else grads


class _SGD:
    def __init__(self, learning_rate=0.01, momentum=0.0, nesterov=False, **kwargs):
        self.learning_rate, self.momentum, self.nesterov = learning_rate, momentum, nesterov
        self._vel = {}

    def apply_gradients(self, grads_and_vars):
        for g, v in grads_and_vars:
            if g is None:
                continue
            if self.momentum:
                vel = self._vel.get(id(v), _torch.zeros_like(v))
                vel = self.momentum * vel - self.learning_rate * g
                self._vel[id(v)] = vel
                step = self.momentum * vel - self.learning_rate * g if self.nesterov else vel
                v.assign_add(step)
            else:
                v.assign_sub(self.learning_rate * g)


class _Optimizers:
    SGD = _SGD


optimizers = _Optimizers()
pi = _math.pi
