"""Device runtime + DeviceArray: the NumPy-interoperable handle the layer mirror passes around.

A DeviceArray owns (or views) dense row-major device memory in fp32 / int32.  Host conversion is
explicit and lazy (`np.asarray(a)`, `float(a)`, indexing); everything else stays on the GPU and
is only *enqueued* on the library stream, so a sequence of layer calls is CUDA-graph capturable."""
import ctypes
import os
import numpy as np
from ._lib import lib, NtError

_state = {"init": False, "device": None, "keep": None}


def init(device=None):
    """Bind this process to one GPU (one process per GPU; LOCAL_RANK picks it under torchrun)."""
    if _state["init"]:
        return _state["device"]
    if device is None:
        device = int(os.environ.get("NTB200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    lib.nt_init(int(device))
    mode = os.environ.get("NTB200_GEMM")
    if mode is not None:
        lib.nt_set_gemm_mode(int(mode))
    _state["init"], _state["device"] = True, device
    return device


def available():
    try:
        n = ctypes.c_int(0)
        lib.load().nt_device_count(ctypes.byref(n))
        return n.value > 0
    except (NtError, OSError):
        return False


class keep_allocations:
    """Context manager: buffers allocated inside are retained in `.buffers` (used while capturing a
    CUDA graph, whose kernels keep writing to those addresses on every replay)."""

    def __enter__(self):
        self.buffers = []
        self._prev = _state["keep"]
        _state["keep"] = self.buffers
        return self

    def __exit__(self, *exc):
        _state["keep"] = self._prev
        return False


def synchronize():
    lib.nt_sync()


def rehome_generation():
    """Bumped whenever parameter tensors move to new device buffers (ParamArena re-homing, fclayer._pad_out).  A
    captured CUDA graph bakes raw parameter addresses in: its owner compares this counter before every replay and
    re-adopts / re-captures when it moved (TrainStep._readopt, KVGenerator._params_moved)."""
    return _state.get("rehome", 0)


def bump_rehome():
    _state["rehome"] = _state.get("rehome", 0) + 1


def launch_count():
    n = ctypes.c_longlong(0)
    lib.nt_launch_count(ctypes.byref(n))
    return n.value


def set_gemm_mode(mode):
    lib.nt_set_gemm_mode(int(mode))
    _state["gemm_mode"] = None


def get_gemm_mode():
    """Cached: the engine choice of every Linear / attention call asks for it."""
    if _state.get("gemm_mode") is None:
        m = ctypes.c_int(0)
        lib.nt_get_gemm_mode(ctypes.byref(m))
        _state["gemm_mode"] = m.value
    return _state["gemm_mode"]


def _size_class(b):
    """The pool's size classes (csrc/runtime.cu:size_class)."""
    if b < 512:
        return 512
    if b < (1 << 20):
        return (b + 511) & ~511
    return (b + (1 << 16) - 1) & ~((1 << 16) - 1)


# Free blocks by size class, kept on the Python side: the launch-bound shapes allocate ~70 small tensors per step, and
# two ctypes round trips per tensor (nt_malloc / nt_free) cost more host time than the kernels they feed.  Same
# stream-ordered reuse rule as the C pool behind it (every consumer runs on the library stream).
_free_blocks = {}
_CACHE_LIMIT = 64 << 20


class _Buffer:
    __slots__ = ("ptr", "nbytes", "_cls")

    def __init__(self, nbytes):
        init()
        self._cls = _size_class(max(int(nbytes), 4))
        cached = _free_blocks.get(self._cls)
        if cached:
            self.ptr = cached.pop()
        else:
            p = ctypes.c_void_p()
            lib.nt_malloc(ctypes.byref(p), self._cls)
            self.ptr = p.value
        self.nbytes = int(nbytes)
        if _state["keep"] is not None:          # graph capture: pin every buffer the graph will touch
            _state["keep"].append(self)

    def __del__(self):
        try:
            if self.ptr:
                if self._cls <= _CACHE_LIMIT:
                    _free_blocks.setdefault(self._cls, []).append(self.ptr)
                else:
                    lib.nt_free(self.ptr)
        except Exception:
            pass


def _prod(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


class DeviceArray:
    """Dense device tensor. dtype is np.float32 or np.int32; `host_dtype` is what np.asarray returns
    (float64 by default, like the reference's fp64 layers)."""
    __array_priority__ = 1000

    def __init__(self, shape, dtype=np.float32, buf=None, offset=0, host_dtype=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        n = _prod(self.shape) * self.dtype.itemsize
        self._buf = buf if buf is not None else _Buffer(n)
        self._off = int(offset)
        self.host_dtype = np.dtype(host_dtype) if host_dtype is not None else (
            np.dtype(np.float64) if self.dtype.kind == "f" else np.dtype(np.int64))
        self._argmax = None
        self._split = None          # (hi, lo) bf16 planes of this tensor emitted by its producer (ops.split_bf16 reuses them)

    # ------------------------------------------------------------------ construction
    @staticmethod
    def from_host(a, host_dtype=None):
        a = np.asarray(a)
        if a.dtype.kind in "iub":
            src = np.ascontiguousarray(a, dtype=np.int32)
            hd = np.int64
        else:
            src = np.ascontiguousarray(a, dtype=np.float32)
            hd = host_dtype if host_dtype is not None else (a.dtype if a.dtype.kind == "f" else np.float64)
        out = DeviceArray(src.shape, src.dtype, host_dtype=hd)
        if src.size:
            lib.nt_h2d(out.ptr, src.ctypes.data, src.nbytes)
        return out

    @staticmethod
    def zeros(shape, dtype=np.float32, host_dtype=None):
        out = DeviceArray(shape, dtype, host_dtype=host_dtype)
        if out.nbytes:
            lib.nt_memset(out.ptr, 0, out.nbytes)
        return out

    def empty_like(self, shape=None):
        return DeviceArray(self.shape if shape is None else shape, self.dtype, host_dtype=self.host_dtype)

    # ------------------------------------------------------------------ properties
    @property
    def ptr(self):
        return self._buf.ptr + self._off

    @property
    def size(self):
        return _prod(self.shape)

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def ndim(self):
        return len(self.shape)

    def __len__(self):
        return self.shape[0]

    # ------------------------------------------------------------------ views / copies
    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(int(s) for s in shape)
        if -1 in shape:
            i = shape.index(-1)
            rest = _prod(shape[:i] + shape[i + 1:])
            shape[i] = self.size // rest if rest else 0
        assert _prod(shape) == self.size, "reshape %s -> %s" % (self.shape, shape)
        out = DeviceArray(shape, self.dtype, buf=self._buf, offset=self._off, host_dtype=self.host_dtype)
        out._split = self._split          # the planes are dense copies in the same element order
        return out

    def view(self, offset_elems, shape):
        return DeviceArray(shape, self.dtype, buf=self._buf,
                           offset=self._off + int(offset_elems) * self.dtype.itemsize, host_dtype=self.host_dtype)

    def copy(self):
        out = self.empty_like()
        if self.nbytes:
            lib.nt_d2d(out.ptr, self.ptr, self.nbytes)
        return out

    def __deepcopy__(self, memo):
        return self.copy()

    def astype(self, dtype, copy=True):
        dtype = np.dtype(dtype)
        if dtype.kind == self.dtype.kind:
            out = self.copy() if copy else self
            if dtype.kind == "f":
                out.host_dtype = dtype
            return out
        return DeviceArray.from_host(self.numpy().astype(dtype))

    def fill_zero(self):
        if self.nbytes:
            lib.nt_memset(self.ptr, 0, self.nbytes)

    def set(self, a):
        """Overwrite with host data (shape-checked by size)."""
        src = np.ascontiguousarray(a, dtype=self.dtype)
        assert src.size == self.size, "set: size mismatch %s vs %s" % (src.shape, self.shape)
        if src.size:
            lib.nt_h2d(self.ptr, src.ctypes.data, src.nbytes)

    # ------------------------------------------------------------------ host conversion
    def numpy(self):
        """Raw device dtype (fp32 / int32) copy on the host. Synchronises."""
        out = np.empty(self.shape, dtype=self.dtype)
        if out.size:
            lib.nt_d2h(out.ctypes.data, self.ptr, out.nbytes)
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.numpy().astype(self.host_dtype, copy=False)
        return a if dtype is None else a.astype(dtype, copy=False)

    def __float__(self):
        return float(self.numpy().reshape(-1)[0])

    def __int__(self):
        return int(self.numpy().reshape(-1)[0])

    def __format__(self, spec):
        return format(float(self), spec) if self.size == 1 else format(np.asarray(self), spec)

    def __repr__(self):
        return "DeviceArray(shape=%s, dtype=%s)" % (self.shape, self.dtype)

    def __getitem__(self, idx):
        return np.asarray(self)[idx]

    def tolist(self):
        return np.asarray(self).tolist()

    def item(self):
        return np.asarray(self).reshape(-1)[0].item()

    # NumPy protocol: a few functions are answered on the device, the rest on a host copy
    def __array_function__(self, func, types, args, kwargs):
        if func is np.argmax and self._argmax is not None and len(args) >= 1 and args[0] is self:
            axis = kwargs.get("axis", args[1] if len(args) > 1 else None)
            if axis in (-1, self.ndim - 1):
                return np.asarray(self._argmax)
        if func is np.reshape and args and args[0] is self:
            shape = args[1] if len(args) > 1 else kwargs.get("shape", kwargs.get("newshape"))
            return self.reshape(shape)
        if func is np.shape:
            return self.shape
        if func is np.ndim:
            return self.ndim
        if func is np.size and len(args) == 1:
            return self.size

        def conv(o):
            if isinstance(o, DeviceArray):
                return np.asarray(o)
            if isinstance(o, (list, tuple)):
                return type(o)(conv(x) for x in o)
            return o
        return func(*conv(args), **{k: conv(v) for k, v in kwargs.items()})

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        ins = [np.asarray(x) if isinstance(x, DeviceArray) else x for x in inputs]
        if "out" in kwargs:
            return NotImplemented
        return getattr(ufunc, method)(*ins, **kwargs)

    # host-side arithmetic for the odd script expression (meanloss += loss, delta * 2, ...)
    def _h(self):
        return float(self) if self.size == 1 and self.ndim == 0 else np.asarray(self)

    def __add__(self, o):
        if isinstance(o, DeviceArray) and o.shape == self.shape and self.dtype == np.float32 and o.dtype == np.float32:
            out = self.empty_like()
            lib.nt_add(self.ptr, o.ptr, out.ptr, self.size)
            return out
        return self._h() + (o._h() if isinstance(o, DeviceArray) else o)

    def __radd__(self, o):
        return o + self._h()

    def __iadd__(self, o):
        if isinstance(o, DeviceArray) and o.shape == self.shape:
            lib.nt_add(self.ptr, o.ptr, self.ptr, self.size)
            return self
        return self.__add__(o)

    def __sub__(self, o):
        return self._h() - (o._h() if isinstance(o, DeviceArray) else o)

    def __rsub__(self, o):
        return o - self._h()

    def __mul__(self, o):
        return self._h() * (o._h() if isinstance(o, DeviceArray) else o)

    def __rmul__(self, o):
        return o * self._h()

    def __truediv__(self, o):
        return self._h() / (o._h() if isinstance(o, DeviceArray) else o)

    def __rtruediv__(self, o):
        return o / self._h()

    def __neg__(self):
        return -self._h()

    def __lt__(self, o):
        return self._h() < o

    def __gt__(self, o):
        return self._h() > o

    def __le__(self, o):
        return self._h() <= o

    def __ge__(self, o):
        return self._h() >= o


def as_device(x, host_dtype=None):
    """ndarray / list / DeviceArray -> DeviceArray (identity for DeviceArray: no copy, no sync)."""
    if isinstance(x, DeviceArray):
        return x
    return DeviceArray.from_host(x, host_dtype=host_dtype)


def ptr_or_null(x):
    return None if x is None else x.ptr


class PinnedBuffer:
    """Page-locked host staging area for the end-to-end path (bench.py e2e, TrainStep)."""

    def __init__(self, shape, dtype):
        init()
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.nbytes = _prod(shape) * self.dtype.itemsize
        p = ctypes.c_void_p()
        lib.nt_host_alloc(ctypes.byref(p), max(self.nbytes, 4))
        self.ptr = p.value
        cbuf = (ctypes.c_byte * max(self.nbytes, 4)).from_address(self.ptr)
        self.array = np.frombuffer(cbuf, dtype=self.dtype, count=_prod(shape)).reshape(self.shape)

    def __del__(self):
        try:
            if self.ptr:
                lib.nt_host_free(self.ptr)
        except Exception:
            pass
