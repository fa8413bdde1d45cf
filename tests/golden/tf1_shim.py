"""A TensorFlow-1.x -> torch (CPU, eager, fp32) shim, just wide enough to EXECUTE the
reference's own function bodies for the part-disentanglement path.

TensorFlow 1.14 cannot be installed in the build container (no cp312 wheel, no network),
so the golden vectors under tests/golden/ are produced by make_golden.py, which reads the
reference's source files at generation time, extracts the functions on the path by name
(ast), and executes them unmodified with `tf` bound to this module.  Nothing from the
reference is copied into the repository.  Each shim op follows the documented semantics
of the TF op of the same name; tensors are torch tensors (autograd supplies tf.gradients).
"""
import ast
import types
import numpy as np
import torch

float32 = torch.float32
int32 = torch.int32
int64 = torch.int64

_DT = {'float32': torch.float32, 'int32': torch.int32, 'int64': torch.int64,
       torch.float32: torch.float32, torch.int32: torch.int32, torch.int64: torch.int64}

TAPS = {}          # side channel: intermediate tensors recorded for the fixtures
_GEN = [None]      # torch.Generator used by random_uniform


class TShape(tuple):
    def as_list(self):
        return [int(v) for v in self]


class TFTensor(torch.Tensor):
    """torch.Tensor with the two TF accessors the reference uses on tensors."""

    @property
    def shape(self):
        return TShape(super().shape)

    def get_shape(self):
        return TShape(super().shape)

    # TF tensors are immutable: `x *= y` rebinds the name to a new tensor (cub/code/nn.py:1586)
    def __imul__(self, other):
        return self * other

    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __getitem__(self, idx):
        # TF strided-slice allows x[..., ::-1]; torch does not -> slice, then flip
        if isinstance(idx, tuple) and any(isinstance(i, builtin_slice) and i.step == -1
                                          for i in idx):
            flips, new = [], []
            for d, i in enumerate(idx):
                if isinstance(i, builtin_slice) and i.step == -1:
                    assert i.start is None and i.stop is None
                    flips.append(d)
                    new.append(builtin_slice(None))
                else:
                    assert isinstance(i, builtin_slice), "only plain slices beside ::-1"
                    new.append(i)
            return torch.flip(super().__getitem__(tuple(new)), dims=flips)
        return super().__getitem__(idx)


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        y = x if dtype is None else x.to(dtype)
    else:
        y = torch.as_tensor(np.asarray(x), dtype=dtype)
        if dtype is None and y.dtype == torch.float64:
            y = y.to(torch.float32)
    return y.as_subclass(TFTensor)


def _shape_arg(shape):
    if isinstance(shape, torch.Tensor):
        return [int(v) for v in shape.tolist()]
    return [int(v) for v in shape]


def set_generator(g):
    _GEN[0] = g


# ------------------------------------------------------------------ creation
def constant(value, dtype=None, shape=None):
    return _t(value, _DT.get(dtype, dtype))


def zeros(shape, dtype=float32):
    return _t(torch.zeros(_shape_arg(shape) if not isinstance(shape, int) else shape,
                          dtype=_DT[dtype]))


def ones(shape, dtype=float32):
    return _t(torch.ones(_shape_arg(shape), dtype=_DT[dtype]))


def ones_like(x):
    return _t(torch.ones_like(x))


def zeros_like(x, dtype=None):
    return _t(torch.zeros_like(x, dtype=_DT.get(dtype, dtype)))


def range(n):  # noqa: A001  (tf.range)
    return _t(torch.arange(int(n), dtype=torch.int32))


def linspace(start, stop, num):
    # TF LinSpace kernel: flat(i) = start + step * i, step = (stop - start) / (num - 1)
    num = int(num)
    start_f, stop_f = torch.tensor(start, dtype=float32), torch.tensor(stop, dtype=float32)
    step = (stop_f - start_f) / torch.tensor(float(num - 1), dtype=float32)
    return _t(start_f + step * torch.arange(num, dtype=float32))


def random_uniform(shape, minval=0.0, maxval=1.0, dtype=float32):
    u = torch.rand(_shape_arg(shape), generator=_GEN[0], dtype=float32)
    return _t(u * (maxval - minval) + minval)


# ------------------------------------------------------------------ shape ops
def shape(x):
    return [int(v) for v in torch.Tensor.size(x)]


def reshape(x, shp):
    return _t(torch.reshape(_t(x), [int(v) for v in shp]))


def expand_dims(x, axis):
    return _t(torch.unsqueeze(_t(x), axis))


def transpose(x, perm=None):
    x = _t(x)
    if perm is None:
        perm = list(reversed(list(builtin_range(x.dim()))))
    return _t(x.permute(*perm).contiguous())


def concat(values, axis):
    return _t(torch.cat([_t(v) for v in values], dim=axis))


def stack(values, axis=0):
    return _t(torch.stack([_t(v) for v in values], dim=axis))


def tile(x, multiples):
    return _t(_t(x).repeat(*[int(m) for m in multiples]))


def split(x, num, axis=0):
    return [_t(c) for c in torch.chunk(x, num, dim=axis)]


def slice(x, begin, size):  # noqa: A001
    idx = []
    for b, s, n in zip(begin, size, torch.Tensor.size(x)):
        idx.append(builtin_slice(b, n if s == -1 else b + s))
    return _t(x[tuple(idx)])


def pad(x, paddings, mode="CONSTANT"):
    assert mode == "CONSTANT"
    flat = []
    for lo, hi in reversed(paddings):
        flat += [lo, hi]
    return _t(torch.nn.functional.pad(x, flat))


def reverse(x, axis):
    return _t(torch.flip(x, dims=list(axis)))


def cast(x, dtype):
    dt = _DT.get(dtype, dtype)
    x = _t(x)
    if x.is_floating_point() and dt in (torch.int32, torch.int64):
        return _t(torch.trunc(x).to(dt))        # tf.cast float->int truncates toward zero
    return _t(x.to(dt))


def to_float(x):
    return cast(x, float32)


# ------------------------------------------------------------------ math
def cos(x): return _t(torch.cos(x))
def sin(x): return _t(torch.sin(x))
def log(x): return _t(torch.log(x))
def exp(x): return _t(torch.exp(x))
def square(x): return _t(x * x)
def floor(x): return _t(torch.floor(x))
def equal(a, b): return _t(torch.eq(a, b))
def stop_gradient(x): return _t(x.detach())


def clip_by_value(x, lo, hi):
    return _t(torch.minimum(torch.maximum(x, _t(lo).to(x.dtype)), _t(hi).to(x.dtype)))


def add_n(values):
    acc = values[0]
    for v in values[1:]:
        acc = acc + v
    return _t(acc)


def matmul(a, b):
    return _t(torch.matmul(_t(a), _t(b)))


# the reference inverts the TPS system (cond 15..460) in fp32 (tf.matrix_inverse,
# transformations.py:228).  With MATRIX_INVERSE_FP64 the shim inverts in float64 and rounds the inverse to fp32:
# make_golden.py writes a second set of TPS fixtures (*_inv64.npz) that isolates how much of the distance between
# the reference's output and the oracle / kernels is the fp32 inverse alone.
MATRIX_INVERSE_FP64 = False


def matrix_inverse(x):
    TAPS.setdefault("matrix_inverse_in", []).append(x.detach().clone())
    if MATRIX_INVERSE_FP64:
        return _t(torch.linalg.inv(x.double()).to(x.dtype))
    return _t(torch.linalg.inv(x))


def einsum(eq, *ops):
    return _t(torch.einsum(eq, *[_t(o) for o in ops]))


def _axes(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else (axis,)


def reduce_sum(x, axis=None, keepdims=False, keep_dims=False):
    ax = _axes(axis)
    return _t(torch.sum(x) if ax is None else torch.sum(x, dim=ax, keepdim=keepdims or keep_dims))


def reduce_mean(x, axis=None, keepdims=False, keep_dims=False):
    ax = _axes(axis)
    return _t(torch.mean(x) if ax is None else torch.mean(x, dim=ax, keepdim=keepdims or keep_dims))


def reduce_max(x, axis=None, keepdims=False, keep_dims=False):
    ax = _axes(axis)
    return _t(torch.amax(x, dim=ax, keepdim=keepdims or keep_dims))


def argmax(x, axis=None):
    return _t(torch.argmax(x, dim=axis))        # int64, first index on ties (like tf.argmax)


def one_hot(indices, depth):
    return _t(torch.nn.functional.one_hot(indices.long(), int(depth)).to(float32))


def gather(params, indices, axis=0):
    assert axis == 0
    return _t(params[indices.long()])


def map_fn(fn, elems, dtype=None):
    if isinstance(elems, (tuple, list)):
        n = torch.Tensor.size(elems[0])[0]
        outs = [fn(tuple(e[i] for e in elems)) for i in builtin_range(n)]
    else:
        outs = [fn(elems[i]) for i in builtin_range(torch.Tensor.size(elems)[0])]
    if isinstance(outs[0], (tuple, list)):
        return tuple(_t(torch.stack([o[j] for o in outs])) for j in builtin_range(len(outs[0])))
    return _t(torch.stack(outs))


def minimum(a, b):
    # TF MinimumGrad routes the gradient to `a` where a <= b, else to `b`
    a, b = _t(a), torch.as_tensor(b, dtype=a.dtype) * torch.ones_like(a)
    return _t(torch.where(a <= b, a, b))


def where(cond, x, y):
    return _t(torch.where(cond, _t(x), _t(y)))


def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32):
    return _t(torch.randn(_shape_arg(shape), generator=_GEN[0], dtype=float32) * stddev + mean)


def _conv2d(input, filter, strides, padding):  # noqa: A002
    # tf.nn.conv2d, NHWC input, HWIO filter, stride 1, SAME (odd kernels: symmetric zero padding);
    # a numpy filter is converted to the input's dtype (convert_to_tensor with preferred dtype)
    assert padding == "SAME" and list(strides) == [1, 1, 1, 1]
    w = torch.as_tensor(np.asarray(filter), dtype=input.dtype) if not isinstance(filter, torch.Tensor) else filter
    kh, kw = w.shape[0], w.shape[1]
    assert kh % 2 == 1 and kw % 2 == 1
    y = torch.nn.functional.conv2d(_t(input).permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=(kh // 2, kw // 2))
    return _t(y.permute(0, 2, 3, 1).contiguous())


def _image_gradients(image):
    # tf.image.image_gradients: forward differences, zero in the last row (dy) / last column (dx)
    image = _t(image)
    dy = torch.cat([image[:, 1:] - image[:, :-1], torch.zeros_like(image[:, :1])], dim=1)
    dx = torch.cat([image[:, :, 1:] - image[:, :, :-1], torch.zeros_like(image[:, :, :1])], dim=2)
    return _t(dy), _t(dx)


def _total_variation(images):
    # tf.image.total_variation for a 4-D batch: sum |dy| + sum |dx| over (1,2,3) -> [batch]
    images = _t(images)
    dy = images[:, 1:] - images[:, :-1]
    dx = images[:, :, 1:] - images[:, :, :-1]
    return _t(dy.abs().sum(dim=(1, 2, 3)) + dx.abs().sum(dim=(1, 2, 3)))


class _XentV2(torch.autograd.Function):
    """tf.nn.softmax_cross_entropy_with_logits_v2 (dim = last axis): loss = sum_k labels*(lse - (x - max));
    registered gradient (nn_grad._SoftmaxCrossEntropyWithLogitsGrad, first order): d/dlogits = g*(softmax - labels)
    -- the fused kernel's `backprop` output, NOT multiplied by sum(labels) -- and d/dlabels = -g*log_softmax."""

    @staticmethod
    def forward(ctx, labels, logits):
        lsm = torch.log_softmax(logits, dim=-1)
        ctx.save_for_backward(labels, lsm)
        return -(labels * lsm).sum(dim=-1)

    @staticmethod
    def backward(ctx, g):
        labels, lsm = ctx.saved_tensors
        g = g.unsqueeze(-1)
        return -g * lsm, g * (torch.exp(lsm) - labels)


def _xent_v2(labels, logits, dim=-1):
    assert dim in (-1, logits.dim() - 1)
    return _t(_XentV2.apply(_t(labels).as_subclass(torch.Tensor), _t(logits).as_subclass(torch.Tensor)))


nn = types.SimpleNamespace(softmax=lambda x, axis=-1: _t(torch.softmax(x, dim=axis)), conv2d=_conv2d,
                           softmax_cross_entropy_with_logits_v2=_xent_v2)
image = types.SimpleNamespace(image_gradients=_image_gradients, total_variation=_total_variation)

import builtins as _b  # noqa: E402
builtin_range = _b.range
builtin_slice = _b.slice


# ------------------------------------------------------------------ source extraction
def load_functions(path, names, namespace):
    """Execute the top-level `def`s called `names` from the reference file at `path`
    inside `namespace` (decorators kept).  The source text never leaves this process."""
    src = open(path).read()
    tree = ast.parse(src)
    def _name(n):
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
            return n.name
        if isinstance(n, ast.Assign) and len(n.targets) == 1 and isinstance(n.targets[0], ast.Name):
            return n.targets[0].id      # module-level constants such as nn.py's `difference1d`
        return None
    wanted = [n for n in tree.body if _name(n) in names]
    found = {_name(n) for n in wanted}
    missing = set(names) - found
    assert not missing, (path, missing)
    mod = ast.Module(body=wanted, type_ignores=[])
    exec(compile(mod, path, "exec"), namespace)
    return namespace
