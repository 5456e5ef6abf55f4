"""Host side of the fused ViT-block kernels (csrc/vit_block.cu, `nt_vit_blocks_fwd/bwd` in include/ntb200.h).

A run of consecutive `attention.attention_layer` objects (attention.py:6-114) is executed as ONE forward launch
and ONE backward launch: every image walks the whole stack inside a single CTA.  The layer objects keep the
reference's observable state (`qkv`, `atg__`/`P`, parameter `*_delta` accumulators, pickle layout); only the
launch granularity changes.  Shapes the kernels do not cover fall back to the per-op path of the layer."""
import ctypes
import os

import numpy as np

from ._lib import lib
from .device import DeviceArray, get_gemm_mode

_PTR_FIELDS = ["ln_g", "ln_b", "wqkv", "bqkv", "ln1_g", "ln1_b", "w0", "b0", "w1", "b1",
               "d_ln_g", "d_ln_b", "d_wqkv", "d_bqkv", "d_ln1_g", "d_ln1_b", "d_w0", "d_b0", "d_w1", "d_b1",
               "qkv", "P", "h", "dgelu", "out", "save", "wfrag"]


class NtVitBlock(ctypes.Structure):
    """Mirror of `struct NtVitBlock` (include/ntb200.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in _PTR_FIELDS]


def enabled():
    return os.environ.get("NTB200_FUSED_VIT", "1") != "0"


def supported(N, D, H):
    """True when the fused kernels cover this shape in the current GEMM mode.  Mode 1 (fp32-class tensor-core path) and
    mode 2 (single-pass, bar 1e-2): the fused kernels' 3-pass products are MORE accurate than mode 2 asks for and, at these
    launch-bound shapes, faster than the per-op single-pass path (0.63 vs 0.49 ms per README step), so mode 2 keeps them.
    Mode 0 (exact fp32 SIMT) does not."""
    if not enabled() or get_gemm_mode() not in (1, 2):
        return False
    if tc_supported(N, D, H):
        return True
    ok = ctypes.c_int(0)
    lib.nt_vit_blocks_supported(int(N), int(D), int(H), ctypes.byref(ok))
    return bool(ok.value)


def tc_supported(N, D, H):
    """The tcgen05 / TMEM kernels (csrc/vit_block_tc.cu) cover this shape (the README configuration: D = 96, 4 heads,
    N <= 64) AND are selected.  They are OPT-IN (NTB200_VIT_TC=1): measured on B200 at B = 100 (tools/vit_tc_time.py,
    DESIGN.md 4.4) the forward is faster than the mma.sync kernels of csrc/vit_block.cu (138 vs 161 us) but the backward is
    slower (310 vs 239 us), so the default stack stays on the older kernels until the backward catches up."""
    if os.environ.get("NTB200_VIT_TC", "0") != "1" or not enabled() or get_gemm_mode() not in (1, 2):
        return False
    ok = ctypes.c_int(0)
    lib.nt_vit_tc_supported(int(N), int(D), int(H), ctypes.byref(ok))
    return bool(ok.value)


_SAVE_BYTES = {}


def _save_bytes(D):
    if D not in _SAVE_BYTES:
        n = ctypes.c_size_t(0)
        lib.nt_vit_blocks_save_bytes(int(D), ctypes.byref(n))
        _SAVE_BYTES[D] = n.value
    return _SAVE_BYTES[D]


def _wfrag(layer, D, tc=False):
    name = "_wimg_tc" if tc else "_wfrag"
    w = getattr(layer, name, None)
    if w is None:
        nbytes = ctypes.c_size_t(0)
        (lib.nt_vit_tc_wimg_bytes if tc else lib.nt_vit_blocks_wfrag_bytes)(int(D), ctypes.byref(nbytes))
        w = DeviceArray((nbytes.value // 4,))
        setattr(layer, name, w)
    return w


def _null_ptr(x):
    return None if x is None else x.ptr


def untile(a, lead, N, W):
    """Host view of a tensor the tcgen05 kernels keep TILED: per leading index [ceil(W/4)][N][4] (the 16-byte unit u of every
    row contiguous over the rows, so a warp's 32 rows store 512 contiguous bytes) -> row-major lead + (N, W)."""
    U = (W + 3) // 4
    a = np.asarray(a).reshape(tuple(lead) + (U, N, 4))
    nd = len(lead)
    a = np.moveaxis(a, nd, nd + 1).reshape(tuple(lead) + (N, U * 4))
    return a[..., :W]


class Tiled:
    """A device tensor in the tiled layout that answers NumPy like its row-major equivalent (np.asarray, shape, ptr).
    Only host code that inspects the layers' saved activations (tests, atg__) ever converts it."""

    def __init__(self, dev, lead, N, W, host_dtype):
        self.dev, self.lead, self.N, self.W, self.host_dtype = dev, tuple(lead), N, W, host_dtype
        self.shape = self.lead + (N, W)

    @property
    def ptr(self):
        return self.dev.ptr

    def __array__(self, dtype=None, copy=None):
        a = untile(self.dev.numpy(), self.lead, self.N, self.W).astype(self.host_dtype, copy=False)
        return a if dtype is None else a.astype(dtype, copy=False)


def _tiled(lead, N, W, hd):
    return Tiled(DeviceArray(tuple(lead) + ((W + 3) // 4, N, 4), host_dtype=hd), lead, N, W, hd)


def prepare(layers):
    """Split the current weights of `layers` into mma-fragment order ahead of forward_stack(..., prepared=True)."""
    D, H = layers[0].embed_dim, layers[0].num_h
    for l in layers:
        _wfrag(l, D)
    arr = _block_array(layers, weights_only=True)
    lib.nt_vit_blocks_prepare(ctypes.addressof(arr), len(layers), D, H)


def _block_array(layers, weights_only=False, tc=False):
    arr = (NtVitBlock * len(layers))()
    for s, l in zip(arr, layers):
        wf = l._wimg_tc if tc else l._wfrag
        if weights_only:
            s.wqkv, s.w0, s.w1, s.wfrag = l.qkvfc.params.ptr, l.fc0.params.ptr, l.fc1.params.ptr, wf.ptr
            continue
        s.ln_g, s.ln_b = l.norm.gamma.ptr, l.norm.beta.ptr
        s.wqkv, s.bqkv = l.qkvfc.params.ptr, l.qkvfc.bias_params.ptr
        s.ln1_g, s.ln1_b = l.norm1.gamma.ptr, l.norm1.beta.ptr
        s.w0, s.b0 = l.fc0.params.ptr, l.fc0.bias_params.ptr
        s.w1, s.b1 = l.fc1.params.ptr, l.fc1.bias_params.ptr
        s.d_ln_g, s.d_ln_b = l.norm.gamma_delta.ptr, l.norm.beta_delta.ptr
        s.d_wqkv, s.d_bqkv = l.qkvfc.params_delta.ptr, l.qkvfc.bias_delta.ptr
        s.d_ln1_g, s.d_ln1_b = l.norm1.gamma_delta.ptr, l.norm1.beta_delta.ptr
        s.d_w0, s.d_b0 = l.fc0.params_delta.ptr, l.fc0.bias_delta.ptr
        s.d_w1, s.d_b1 = l.fc1.params_delta.ptr, l.fc1.bias_delta.ptr
        s.qkv, s.P, s.h, s.dgelu, s.out = l.qkv.ptr, _null_ptr(l.P), l.input1.ptr, l._dgelu.ptr, l._out.ptr
        s.save = _null_ptr(l._save)
        s.wfrag = wf.ptr
    return arr


def forward_stack(layers, x, prepared=False, want_P=True):
    """attention_layer.forward (attention.py:20-55) of every layer in `layers`, chained, in one launch.
    want_P=False (TrainStep): the tcgen05 kernels do not materialise the softmax matrices (the reference's atg__); the
    backward recomputes them from q and k."""
    B, N, D = x.shape
    H = layers[0].num_h
    hd = x.host_dtype
    tc = tc_supported(N, D, H)
    for l in layers:
        assert l.embed_dim == D and l.num_h == H, "fused stack needs identical blocks"
        l.batch, l.block = B, N
        l._mask_kind = 0
        last = l is layers[-1]
        if tc:          # saved activations in the kernels' tiled layout (see Tiled); the stack's output stays row-major
            l.qkv = _tiled((B,), N, 3 * D, hd)
            l.P = _tiled((B, H), N, N, hd) if want_P else None
            l.input1 = _tiled((B,), N, D, hd)
            l._dgelu = _tiled((B,), N, 2 * D, hd)                     # the fc0 PRE-activation
            l._save = None
            l._out = DeviceArray((B, N, D), host_dtype=hd) if last else _tiled((B,), N, D, hd)
        else:
            l.qkv = DeviceArray((B, N, 3 * D), host_dtype=hd)
            l.P = DeviceArray((B, H, N, N), host_dtype=hd)
            l.input1 = DeviceArray((B, N, D), host_dtype=hd)
            l._dgelu = DeviceArray((B, N, 2 * D), host_dtype=hd)
            l._save = DeviceArray((B, _save_bytes(D) // 4))
            l._out = DeviceArray((B, N, D), host_dtype=hd)
        l.out = l.out1 = l.out2 = None          # LN outputs / GELU output are recomputed in backward, not stored
        l._fused = "tc" if tc else "mma"
        _wfrag(l, D, tc)
    layers[0]._x_in = x
    for prev, l in zip(layers, layers[1:]):
        l._x_in = prev._out
    arr = _block_array(layers, tc=tc)
    if tc:
        lib.nt_vit_tc_fwd(ctypes.addressof(arr), len(layers), x.ptr, B, N, D, H, int(bool(prepared)))
    else:
        lib.nt_vit_blocks_fwd(ctypes.addressof(arr), len(layers), x.ptr, B, N, D, H, int(bool(prepared)))
    return layers[-1]._out


def backward_stack(layers, delta):
    """attention_layer.backward (attention.py:57-89) of the stack, last layer first, in one launch."""
    x = layers[0]._x_in
    B, N, D = x.shape
    H = layers[0].num_h
    dx = DeviceArray((B, N, D), host_dtype=delta.host_dtype)
    if layers[0]._fused == "tc":
        arr = _block_array(layers, tc=True)
        lib.nt_vit_tc_bwd(ctypes.addressof(arr), len(layers), x.ptr, delta.ptr, dx.ptr, B, N, D, H)
        return dx
    scratch = DeviceArray((2, B, N, D))
    arr = _block_array(layers)
    lib.nt_vit_blocks_bwd(ctypes.addressof(arr), len(layers), x.ptr, delta.ptr, dx.ptr, scratch.ptr, B, N, D, H)
    return dx


def fusable_run(layers, start, block_type, x_shape):
    """Length of the run of fusable `block_type` layers beginning at layers[start] (0 if the first is not)."""
    if len(x_shape) != 3:
        return 0
    _, N, D = x_shape
    n = 0
    while start + n < len(layers) and n < 16:
        l = layers[start + n]
        if type(l) is not block_type or l.embed_dim != D or l.num_h != layers[start].num_h:
            break
        n += 1
    if n and not supported(N, D, layers[start].num_h):
        return 0
    return n
