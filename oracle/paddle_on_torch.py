"""A minimal PaddlePaddle (2.1-2.4 dygraph) API surface implemented on torch-CPU.  TEST INFRASTRUCTURE ONLY.

Purpose: execute the REFERENCE'S OWN, UNMODIFIED Python sources (imported from /root/reference by
``oracle/run_reference.py``) in this container, where PaddlePaddle itself cannot be installed, so that golden
vectors for the hot path come from the reference's code — its reshapes, transposes, splits, level loops, window
loops and parameter layouts — rather than from our restatement of it.  Only the primitive operators are mapped to
torch; every mapping is one line and states the Paddle semantics it assumes (from the Paddle 2.x API docs):

  * ``nn.Linear``: ``y = x @ W + b`` with ``W`` stored ``[in, out]``; ``F.linear(x, weight, bias)`` likewise.
  * ``nn.Conv2D`` weight ``[out, in/groups, kh, kw]`` (same as torch); ``bias_attr=False`` -> no bias.
  * ``nn.LayerNorm`` / ``nn.GroupNorm`` / ``nn.BatchNorm2D``: epsilon 1e-5, affine ``weight`` / ``bias``.
  * ``nn.GELU``: exact erf form (``approximate=False``); ``nn.Dropout``: identity in eval mode.
  * ``F.softmax(x, axis=-1)``; ``F.grid_sample``: grid last dim (x, y) in [-1, 1], same modes as torch.
  * ``F.interpolate``: default mode 'nearest'; bilinear with ``align_corners=False`` uses half-pixel centres
    (``align_mode=0`` default) = torch's behaviour; with ``align_corners=True`` = torch's.
  * ``Tensor.transpose(perm)`` is a permutation; ``Tensor.max(axis, keepdim)`` returns values only;
    ``Tensor.split(num_or_sections, axis)``: an int is the NUMBER of sections, a list gives section sizes;
    ``Tensor.shape`` is a python list; ``reshape`` treats 0 as "copy this dim"; ``Tensor.numpy()`` of a scalar
    and reductions without an axis (``paddle.sum(x)``) have shape ``[1]`` (Paddle < 2.5 has no 0-d tensors).
  * ``paddle.meshgrid`` uses 'ij' indexing; ``paddle.argmax`` returns the first maximal index.

Anything the hot path's sources do not touch is absent on purpose; an AttributeError from this module means the
reference used an API that has not been mapped (map it here, citing the Paddle doc semantics).
"""
from __future__ import annotations

import math
import sys
import types
from typing import Sequence

import numpy as np
import torch
import torch.nn.functional as TF
import torch.utils.dlpack

_T = torch.Tensor
_DEVICE = [torch.device("cpu")]          # paddle.set_device("gpu") moves creation ops (and new parameters) to cuda:0


def set_device(name):
    name = str(name)
    _DEVICE[0] = torch.device("cuda", int(name.split(":")[1]) if ":" in name else 0) if name.startswith("gpu") \
        else torch.device("cpu")
    return _DEVICE[0]


def get_device():
    d = _DEVICE[0]
    return "cpu" if d.type == "cpu" else f"gpu:{d.index or 0}"


class _Place:
    def __init__(self, dev):
        self._dev = dev

    def is_gpu_place(self):
        return self._dev.type == "cuda"

    def is_cpu_place(self):
        return self._dev.type == "cpu"

_DTYPES = {
    "float32": torch.float32, "float64": torch.float64, "float16": torch.float16, "bfloat16": torch.bfloat16,
    "int32": torch.int32, "int64": torch.int64, "bool": torch.bool, "uint8": torch.uint8, "int8": torch.int8,
}


def _dt(d):
    if d is None or isinstance(d, torch.dtype):
        return d
    return _DTYPES[str(d).replace("paddle.", "")]


def _raw(x):
    """paddle Tensor -> plain torch.Tensor (python scalars / lists pass through)."""
    if isinstance(x, torch.Tensor):
        return x.as_subclass(torch.Tensor) if type(x) is not torch.Tensor else x
    return x


def _ints(seq):
    return [int(_raw(s)) for s in seq]


def _wrap(x):
    if isinstance(x, torch.Tensor) and not isinstance(x, Tensor):
        return x.as_subclass(Tensor)
    if isinstance(x, (list, tuple)):
        return type(x)(_wrap(t) for t in x)
    return x


class Tensor(torch.Tensor):
    """torch.Tensor with the Paddle method semantics the reference relies on."""

    @property
    def shape(self):  # paddle: python list
        return list(_T.size(self))

    @property
    def place(self):
        return _Place(_T.device.__get__(self))

    @property
    def stop_gradient(self):
        return not _T.requires_grad.__get__(self)

    @stop_gradient.setter
    def stop_gradient(self, v):
        _T.requires_grad_(self, not v)

    @property
    def inplace_version(self):      # paddle.Tensor.inplace_version: bumped by every in-place write (set_value, optimiser)
        return _T._version.__get__(self)

    def numpy(self):
        a = _raw(self).detach().cpu().numpy()
        return a.reshape(1) if a.ndim == 0 else a          # Paddle < 2.5: scalars are shape [1]

    def astype(self, dtype):
        return _wrap(_raw(self).to(_dt(dtype)))

    def transpose(self, *perm, **kw):
        if "perm" in kw:
            perm = (kw["perm"],)
        if len(perm) == 1 and isinstance(perm[0], (list, tuple)):
            return _wrap(_raw(self).permute(*_ints(perm[0])))
        return _wrap(_T.transpose(_raw(self), *perm))         # torch-style call from torch internals

    def reshape(self, *shape, **kw):
        if "shape" in kw:
            shape = (kw["shape"],)
        if len(shape) == 1 and isinstance(shape[0], (list, tuple)):
            shape = shape[0]
        shape = _ints(shape)
        cur = list(_T.size(self))
        shape = [cur[i] if s == 0 else s for i, s in enumerate(shape)]          # 0 = copy the input's dim
        return _wrap(_raw(self).reshape(shape))

    def flatten(self, start_axis=0, stop_axis=-1):
        return _wrap(_raw(self).flatten(start_axis, stop_axis))

    def split(self, num_or_sections, axis=0):
        r = _raw(self)
        if isinstance(num_or_sections, int):
            assert r.shape[axis] % num_or_sections == 0
            return _wrap(list(torch.split(r, r.shape[axis] // num_or_sections, dim=axis)))
        return _wrap(list(torch.split(r, _ints(num_or_sections), dim=axis)))

    def flip(self, axis):
        return _wrap(torch.flip(_raw(self), [axis] if isinstance(axis, int) else list(axis)))

    def tile(self, *reps):
        if len(reps) == 1 and isinstance(reps[0], (list, tuple)):
            reps = reps[0]
        return _wrap(_raw(self).repeat(*_ints(reps)) if len(reps) >= _raw(self).dim() else torch.tile(_raw(self), _ints(reps)))

    def _reduce(self, fn, axis, keepdim):
        r = _raw(self)
        if axis is None:
            return _wrap(fn(r).reshape(1))                    # Paddle < 2.5: full reductions have shape [1]
        return _wrap(fn(r, dim=axis, keepdim=keepdim))

    def max(self, axis=None, keepdim=False):
        return self._reduce(torch.amax, axis, keepdim)

    def min(self, axis=None, keepdim=False):
        return self._reduce(torch.amin, axis, keepdim)

    def sum(self, axis=None, dtype=None, keepdim=False):
        r = _raw(self)
        if dtype is not None:
            r = r.to(_dt(dtype))
        if axis is None:
            return _wrap(r.sum().reshape(1))
        return _wrap(r.sum(dim=axis, keepdim=keepdim))

    def prod(self, axis=None, keepdim=False):
        r = _raw(self)
        return _wrap(r.prod().reshape(1) if axis is None else r.prod(dim=axis, keepdim=keepdim))

    def mean(self, axis=None, keepdim=False):
        r = _raw(self)
        return _wrap(r.mean().reshape(1) if axis is None else r.mean(dim=axis, keepdim=keepdim))

    def cumsum(self, axis=None, dtype=None):
        r = _raw(self)
        return _wrap(torch.cumsum(r.flatten() if axis is None else r, dim=0 if axis is None else axis, dtype=_dt(dtype)))

    def unsqueeze(self, axis):
        return _wrap(_raw(self).unsqueeze(axis))

    def squeeze(self, axis=None):
        r = _raw(self)
        return _wrap(r.squeeze() if axis is None else r.squeeze(axis))

    def clip(self, min=None, max=None):
        return _wrap(torch.clamp(_raw(self), min=min, max=max))

    def unbind(self, axis=0):
        return _wrap(list(torch.unbind(_raw(self), dim=axis)))

    def set_value(self, value):
        with torch.no_grad():
            _raw(self).copy_(torch.as_tensor(_raw(value) if isinstance(value, torch.Tensor) else np.asarray(value)).reshape(
                _T.size(self)))

    def dim(self):
        return _T.dim(self)

    def __deepcopy__(self, memo):                            # copy.deepcopy of Layers (utils.py:31 _get_clones)
        t = _raw(self).detach().clone().as_subclass(Tensor)
        if getattr(self, "_is_param", False):
            t = torch.nn.Parameter(t, requires_grad=self.requires_grad)
        memo[id(self)] = t
        return t

    def __int__(self):
        return int(_raw(self).reshape(-1)[0]) if _T.numel(self) == 1 else _T.__int__(self)


def to_tensor(data, dtype=None, place=None, stop_gradient=True):
    if isinstance(data, torch.Tensor):
        t = _raw(data).clone()
    else:
        a = np.asarray(data)
        if dtype is None and a.dtype == np.float64:
            a = a.astype(np.float32)                            # paddle default float dtype
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(_dt(dtype))
    return _wrap(t.to(_DEVICE[0]))


def _shape_arg(shape):
    return _ints(shape) if isinstance(shape, (list, tuple)) else [int(shape)]


def _fdt(dtype):
    return _dt(dtype) if dtype is not None else torch.float32


def zeros(shape, dtype=None):
    return _wrap(torch.zeros(_shape_arg(shape), dtype=_fdt(dtype), device=_DEVICE[0]))


def ones(shape, dtype=None):
    return _wrap(torch.ones(_shape_arg(shape), dtype=_fdt(dtype), device=_DEVICE[0]))


def empty(shape, dtype=None):
    return _wrap(torch.empty(_shape_arg(shape), dtype=_fdt(dtype), device=_DEVICE[0]))


def zeros_like(x, dtype=None):
    return _wrap(torch.zeros_like(_raw(x), dtype=_dt(dtype)))


def is_grad_enabled():
    return torch.is_grad_enabled()


def full_like(x, fill_value, dtype=None):
    return _wrap(torch.full_like(_raw(x), fill_value, dtype=_dt(dtype)))


def arange(start=0, end=None, step=1, dtype=None):
    if end is None:
        start, end = 0, start
    all_int = all(isinstance(v, (int, np.integer)) for v in (start, end, step))
    return _wrap(torch.arange(start, end, step, device=_DEVICE[0],
                              dtype=_dt(dtype) if dtype is not None else (torch.int64 if all_int else torch.float32)))


def linspace(start, stop, num, dtype=None):
    return _wrap(torch.linspace(float(start), float(stop), int(num), dtype=_fdt(dtype), device=_DEVICE[0]))


def meshgrid(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = args[0]
    return _wrap(list(torch.meshgrid(*[_raw(a) for a in args], indexing="ij")))


def stack(x, axis=0):
    return _wrap(torch.stack([_raw(t) for t in x], dim=axis))


def concat(x, axis=0):
    return _wrap(torch.cat([_raw(t) for t in x], dim=axis))


def split(x, num_or_sections, axis=0):
    return _wrap(x).split(num_or_sections, axis)


def reshape(x, shape):
    return _wrap(x).reshape(shape)


def transpose(x, perm):
    return _wrap(x).transpose(perm)


def squeeze(x, axis=None):
    return _wrap(x).squeeze(axis)


def sum(x, axis=None, dtype=None, keepdim=False):  # noqa: A001 (mirrors paddle.sum)
    return _wrap(x).sum(axis, dtype, keepdim)


def matmul(x, y, transpose_x=False, transpose_y=False):
    a, b = _raw(x), _raw(y)
    if transpose_x:
        a = a.transpose(-1, -2)
    if transpose_y:
        b = b.transpose(-1, -2)
    return _wrap(torch.matmul(a, b))


def argmax(x, axis=None, keepdim=False, dtype="int64"):
    r = _raw(x)
    out = torch.argmax(r.flatten() if axis is None else r, dim=0 if axis is None else axis, keepdim=keepdim)
    return _wrap(out.to(_dt(dtype)))


def log(x):
    return _wrap(torch.log(_raw(x)))


def uniform(shape, dtype=None, min=-1.0, max=1.0, seed=0):
    return _wrap(torch.empty(_shape_arg(shape), dtype=_fdt(dtype), device=_DEVICE[0]).uniform_(min, max))


def normal(mean=0.0, std=1.0, shape=None):
    return _wrap(torch.empty(_shape_arg(shape), dtype=torch.float32, device=_DEVICE[0]).normal_(mean, std))


def rand(shape, dtype=None):
    return _wrap(torch.rand(_shape_arg(shape), dtype=_fdt(dtype), device=_DEVICE[0]))


class ParamAttr:
    def __init__(self, name=None, initializer=None, learning_rate=1.0, regularizer=None, trainable=True, **kw):
        self.learning_rate = learning_rate
        self.initializer = initializer


def _param(shape, fill=None):
    t = torch.empty(list(shape), dtype=torch.float32, device=_DEVICE[0])
    if fill is None:
        bound = math.sqrt(6.0 / max(1, (shape[0] + shape[-1]))) if len(shape) >= 2 else 0.0
        t.uniform_(-bound, bound) if bound else t.zero_()
    else:
        t.fill_(fill)
    return torch.nn.Parameter(t.as_subclass(Tensor), requires_grad=True)      # paddle: stop_gradient=False


# ---------------------------------------------------------------------------------------------------------------------
# paddle.nn
# ---------------------------------------------------------------------------------------------------------------------
class Layer(torch.nn.Module):
    _dtype = "float32"

    def __call__(self, *a, **kw):
        return _wrap(super().__call__(*a, **kw))

    def create_parameter(self, shape, attr=None, dtype=None, is_bias=False, default_initializer=None):
        return _param(shape, 0.0 if is_bias else None)

    def sublayers(self, include_self=False):
        mods = list(self.modules())
        return mods if include_self else mods[1:]

    def add_sublayer(self, name, sublayer):
        self.add_module(name, sublayer)
        return sublayer

    def set_state_dict(self, sd):
        own = dict(self.named_parameters())
        own.update(dict(self.named_buffers()))
        missing = [k for k in own if k not in sd]
        if missing:
            raise KeyError(f"missing keys: {missing[:5]} ...")
        with torch.no_grad():
            for k, p in own.items():
                _raw(p).copy_(torch.as_tensor(np.asarray(sd[k])).reshape(p.shape))


class Linear(Layer):
    def __init__(self, in_features, out_features, weight_attr=None, bias_attr=None, name=None):
        super().__init__()
        self.weight = _param([in_features, out_features])                        # paddle layout [in, out]
        self.bias = None if bias_attr is False else _param([out_features], 0.0)

    def forward(self, x):
        return functional.linear(x, self.weight, self.bias)


class Conv2D(Layer):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 padding_mode="zeros", weight_attr=None, bias_attr=None, data_format="NCHW"):
        super().__init__()
        k = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.weight = _param([out_channels, in_channels // groups, *k])
        self.bias = None if bias_attr is False else _param([out_channels], 0.0)
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups

    def forward(self, x):
        return _wrap(TF.conv2d(_raw(x), _raw(self.weight), None if self.bias is None else _raw(self.bias), self.stride,
                               self.padding, self.dilation, self.groups))


class GroupNorm(Layer):
    def __init__(self, num_groups, num_channels, epsilon=1e-5, weight_attr=None, bias_attr=None, data_format="NCHW"):
        super().__init__()
        self.num_groups, self.epsilon = num_groups, epsilon
        self.weight, self.bias = _param([num_channels], 1.0), _param([num_channels], 0.0)

    def forward(self, x):
        return _wrap(TF.group_norm(_raw(x), self.num_groups, _raw(self.weight), _raw(self.bias), self.epsilon))


class LayerNorm(Layer):
    def __init__(self, normalized_shape, epsilon=1e-5, weight_attr=None, bias_attr=None):
        super().__init__()
        self.shape_ = [normalized_shape] if isinstance(normalized_shape, int) else list(normalized_shape)
        self.epsilon = epsilon
        self.weight, self.bias = _param(self.shape_, 1.0), _param(self.shape_, 0.0)

    def forward(self, x):
        return _wrap(TF.layer_norm(_raw(x), self.shape_, _raw(self.weight), _raw(self.bias), self.epsilon))


class BatchNorm2D(Layer):
    def __init__(self, num_features, momentum=0.9, epsilon=1e-5, weight_attr=None, bias_attr=None, data_format="NCHW"):
        super().__init__()
        self.epsilon = epsilon
        self.weight, self.bias = _param([num_features], 1.0), _param([num_features], 0.0)
        self.register_buffer("_mean", torch.zeros(num_features, device=_DEVICE[0]))
        self.register_buffer("_variance", torch.ones(num_features, device=_DEVICE[0]))

    def forward(self, x):
        assert not self.training, "the shim runs BatchNorm in eval mode only"
        return _wrap(TF.batch_norm(_raw(x), self._mean, self._variance, _raw(self.weight), _raw(self.bias), False, 0.0,
                                   self.epsilon))


SyncBatchNorm = BatchNorm2D


class Embedding(Layer):
    def __init__(self, num_embeddings, embedding_dim, padding_idx=None, sparse=False, weight_attr=None):
        super().__init__()
        self.weight = _param([num_embeddings, embedding_dim])

    def forward(self, x):
        return _wrap(TF.embedding(_raw(x), _raw(self.weight)))


class Dropout(Layer):
    def __init__(self, p=0.5, axis=None, mode="upscale_in_train"):
        super().__init__()
        self.p = p

    def forward(self, x):
        assert not self.training or self.p == 0, "the shim runs Dropout in eval mode only"
        return x


class GELU(Layer):
    def __init__(self, approximate=False):
        super().__init__()
        self.approximate = approximate

    def forward(self, x):
        return _wrap(TF.gelu(_raw(x), approximate="tanh" if self.approximate else "none"))


class ReLU(Layer):
    def __init__(self, *a, **kw):
        super().__init__()

    def forward(self, x):
        return _wrap(TF.relu(_raw(x)))


class Identity(Layer):
    def forward(self, x):
        return x


class Sequential(Layer):
    def __init__(self, *layers):
        super().__init__()
        for i, l in enumerate(layers):
            self.add_module(str(i), l)

    def __getitem__(self, i):
        return list(self.children())[i]

    def __len__(self):
        return len(list(self.children()))

    def forward(self, x):
        for l in self.children():
            x = l(x)
        return x


class LayerList(Layer):
    def __init__(self, sublayers=None):
        super().__init__()
        for l in (sublayers or []):
            self.append(l)

    def append(self, l):
        self.add_module(str(len(self._modules)), l)
        return self

    def __getitem__(self, i):
        return list(self.children())[i]

    def __iter__(self):
        return iter(list(self.children()))

    def __len__(self):
        return len(self._modules)


# ---------------------------------------------------------------------------------------------------------------------
# paddle.nn.functional
# ---------------------------------------------------------------------------------------------------------------------
class _Functional(types.ModuleType):
    @staticmethod
    def linear(x, weight, bias=None, name=None):
        y = torch.matmul(_raw(x), _raw(weight))                                 # weight [in, out]
        return _wrap(y if bias is None else y + _raw(bias))

    @staticmethod
    def softmax(x, axis=-1, dtype=None, name=None):
        return _wrap(TF.softmax(_raw(x), dim=axis, dtype=_dt(dtype)))

    @staticmethod
    def sigmoid(x, name=None):
        return _wrap(torch.sigmoid(_raw(x)))

    @staticmethod
    def relu(x, name=None):
        return _wrap(TF.relu(_raw(x)))

    @staticmethod
    def gelu(x, approximate=False, name=None):
        return _wrap(TF.gelu(_raw(x), approximate="tanh" if approximate else "none"))

    @staticmethod
    def dropout(x, p=0.5, axis=None, training=True, mode="upscale_in_train", name=None):
        assert not training or p == 0, "the shim runs dropout in eval mode only"
        return x

    @staticmethod
    def grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=True, name=None):
        return _wrap(TF.grid_sample(_raw(x), _raw(grid), mode=mode, padding_mode=padding_mode, align_corners=align_corners))

    @staticmethod
    def interpolate(x, size=None, scale_factor=None, mode="nearest", align_corners=False, align_mode=0,
                    data_format="NCHW", name=None):
        r = _raw(x)
        if size is not None:
            size = _ints(size) if isinstance(size, (list, tuple)) else _ints(_raw(size).tolist())
        if mode == "nearest":
            return _wrap(TF.interpolate(r, size=size, scale_factor=scale_factor, mode="nearest"))
        assert align_corners or align_mode == 0, "align_mode=1 (no half-pixel shift) is not mapped"
        return _wrap(TF.interpolate(r, size=size, scale_factor=scale_factor, mode=mode, align_corners=align_corners))

    @staticmethod
    def one_hot(x, num_classes, name=None):
        return _wrap(TF.one_hot(_raw(x).long(), int(num_classes)).to(torch.float32))     # paddle: float32 output


functional = _Functional("paddle.nn.functional")


def _convert_attention_mask(attn_mask, dtype):
    m = _raw(attn_mask)
    if m.dtype == torch.bool:
        return _wrap((m.to(dtype) - 1.0) * 1e9)
    if not m.dtype.is_floating_point:
        return _wrap((m.to(dtype) - 1.0) * 1e9)
    return _wrap(m.to(dtype))


class _NoGrad:
    def __enter__(self):
        self._g = torch.no_grad()
        self._g.__enter__()

    def __exit__(self, *a):
        return self._g.__exit__(*a)

    def __call__(self, fn=None):
        if fn is None:
            return _NoGrad()
        return torch.no_grad()(fn)


class PyLayer:
    """paddle.autograd.PyLayer on torch.autograd.Function: ``forward(ctx, *args)`` / ``backward(ctx, *grads)`` static
    methods, ``ctx.save_for_backward`` / ``ctx.saved_tensor()``; backward returns one gradient per TENSOR input."""

    @classmethod
    def apply(cls, *args):
        is_t = [isinstance(a, torch.Tensor) for a in args]

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *a):
                ctx.saved_tensor = lambda: tuple(_wrap(t) for t in ctx.saved_tensors)
                out = cls.forward(ctx, *[_wrap(x) for x in a])
                return tuple(_raw(o) for o in out) if isinstance(out, (tuple, list)) else _raw(out)

            @staticmethod
            def backward(ctx, *grads):
                res = cls.backward(ctx, *[_wrap(g) for g in grads])
                res = list(res) if isinstance(res, (tuple, list)) else [res]
                assert len(res) == len([t for t in is_t if t]), "PyLayer.backward must return one gradient per tensor input"
                it = iter(res)
                return tuple((_raw(next(it)) if t else None) for t in is_t)

        return _wrap(_Fn.apply(*[_raw(a) for a in args]))


class _CudaStream:
    @property
    def cuda_stream(self):
        return torch.cuda.current_stream().cuda_stream


class _Init:
    def __init__(self, *a, **kw):
        pass


def install():
    """Registers the shim as ``paddle`` (and sub-modules) in sys.modules.  Refuses to shadow a real PaddlePaddle."""
    if "paddle" in sys.modules:
        if getattr(sys.modules["paddle"], "__emrt_shim__", False):
            return sys.modules["paddle"]
        raise RuntimeError("a real `paddle` is already imported; use it instead of the shim")
    me = sys.modules[__name__]
    paddle = types.ModuleType("paddle")
    paddle.__emrt_shim__ = True
    for name in ("set_device", "get_device", "empty", "zeros_like", "is_grad_enabled"):
        setattr(paddle, name, getattr(me, name))
    autograd = types.ModuleType("paddle.autograd")
    autograd.PyLayer = PyLayer
    device = types.ModuleType("paddle.device")
    device_cuda = types.ModuleType("paddle.device.cuda")
    device_cuda.current_stream = lambda *a: _CudaStream()
    device.cuda, device.set_device, device.get_device = device_cuda, set_device, get_device
    utils = types.ModuleType("paddle.utils")
    dl = types.ModuleType("paddle.utils.dlpack")
    dl.to_dlpack = lambda t: torch.utils.dlpack.to_dlpack(_raw(t).detach())
    dl.from_dlpack = lambda cap: _wrap(torch.utils.dlpack.from_dlpack(cap))
    utils.dlpack = dl
    paddle.utils = utils
    sys.modules.update({"paddle.utils": utils, "paddle.utils.dlpack": dl})
    paddle.autograd, paddle.device = autograd, device
    sys.modules.update({"paddle.autograd": autograd, "paddle.device": device, "paddle.device.cuda": device_cuda})
    for name in ("Tensor", "to_tensor", "zeros", "ones", "full_like", "arange", "linspace", "meshgrid", "stack",
                 "concat", "split", "reshape", "transpose", "squeeze", "sum", "matmul", "argmax", "log", "uniform",
                 "normal", "rand", "ParamAttr"):
        setattr(paddle, name, getattr(me, name))
    for k, v in _DTYPES.items():
        setattr(paddle, k, v)
    paddle.no_grad = _NoGrad()
    paddle.Layer = Layer
    nn = types.ModuleType("paddle.nn")
    for name in ("Layer", "Linear", "Conv2D", "GroupNorm", "LayerNorm", "BatchNorm2D", "SyncBatchNorm", "Embedding",
                 "Dropout", "GELU", "ReLU", "Identity", "Sequential", "LayerList"):
        setattr(nn, name, getattr(me, name))
    nn.functional = functional
    init = types.ModuleType("paddle.nn.initializer")
    for name in ("Normal", "Constant", "XavierUniform", "KaimingNormal", "KaimingUniform", "TruncatedNormal", "Uniform"):
        setattr(init, name, _Init)
    nn.initializer = init
    layer = types.ModuleType("paddle.nn.layer")
    transformer = types.ModuleType("paddle.nn.layer.transformer")
    transformer._convert_attention_mask = _convert_attention_mask
    layer.transformer = transformer
    nn.layer = layer
    reg = types.ModuleType("paddle.regularizer")
    reg.L2Decay = _Init
    vision = types.ModuleType("paddle.vision")
    vops = types.ModuleType("paddle.vision.ops")
    vops.DeformConv2D = type("DeformConv2D", (Layer,), {})
    vision.ops = vops
    paddle.nn, paddle.regularizer, paddle.vision = nn, reg, vision
    sys.modules.update({
        "paddle": paddle, "paddle.nn": nn, "paddle.nn.functional": functional, "paddle.nn.initializer": init,
        "paddle.nn.layer": layer, "paddle.nn.layer.transformer": transformer, "paddle.regularizer": reg,
        "paddle.vision": vision, "paddle.vision.ops": vops,
    })
    return paddle
