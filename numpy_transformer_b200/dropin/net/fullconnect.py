"""net.fullconnect.fclayer on the B200 backend (mirrors the reference's net/fullconnect.py:36-164).
Same constructor, attributes (params, bias_params, params_delta, bias_delta, t) and methods; the
tensors live on the GPU as DeviceArrays and every method only enqueues CUDA work."""
import numpy as np
from numpy_transformer_b200.device import DeviceArray, as_device
from numpy_transformer_b200 import ops
from numpy_transformer_b200._lib import lib


class fclayer(object):
    def __init__(self, infeature, outfeature, bias=False, params=[], bias_params=[], adam=False, float32=False,
                 float16=False, name='', init=''):
        self.infeature, self.outfeature, self.bias = infeature, outfeature, bias
        self.float32, self.float16 = float32, float16
        self.host_dtype = np.float32 if float32 else (np.float16 if float16 else np.float64)
        # same draws, in the same order, as net/fullconnect.py:41-59 (bias is drawn even if unused)
        if len(params):
            w = np.asarray(params)
        else:
            r = np.sqrt(1 / infeature)
            w = np.random.uniform(-r, r, (infeature, outfeature))
        if bias and len(bias_params):
            b = np.asarray(bias_params)
        else:
            r = np.sqrt(1 / infeature)
            b = np.random.uniform(-r, r, (outfeature))
        self.params = DeviceArray.from_host(w.reshape(infeature, outfeature), host_dtype=self.host_dtype)
        self.bias_params = DeviceArray.from_host(b.reshape(outfeature), host_dtype=self.host_dtype)
        self.params_delta = DeviceArray.zeros((infeature, outfeature), host_dtype=self.host_dtype)
        self.bias_delta = DeviceArray.zeros((outfeature,), host_dtype=self.host_dtype)
        self.adam = adam
        self.beta1, self.beta2, self.epsadam = 0.9, 0.999, 10 ** (2 - 10)
        self.moment_p = self.rmsprop_p = self.moment_b = self.rmsprop_b = None
        if adam:
            self.moment_p = DeviceArray.zeros((infeature, outfeature), host_dtype=self.host_dtype)
            self.rmsprop_p = DeviceArray.zeros((infeature, outfeature), host_dtype=self.host_dtype)
            self.moment_b = DeviceArray.zeros((outfeature,), host_dtype=self.host_dtype)
            self.rmsprop_b = DeviceArray.zeros((outfeature,), host_dtype=self.host_dtype)
        self.t = 1
        self.inputs = None

    # --- extended forms used by the fused composite layers
    def _forward(self, x, act=ops.ACT_NONE, want_pre=False, residual=None, split_out=False):
        self.inputs = x
        keep, keep_w = [], []
        out = ops.linear_fwd(x, self.params, self.bias_params if self.bias else None, act, want_pre, residual, keep_xt=keep,
                             split_out=split_out, keep_w=keep_w)
        self._xt = (x, keep[0]) if keep else None       # BF16x3 path: transposed split of x, made in the forward's pass
        self._wsplit = keep_w[0] if keep_w else None    # BF16x3 path: split of the weights for dX (valid until update)
        return out

    def _backward(self, dy, x, act=ops.ACT_NONE, pre=None, addend=None, want_dx=True, split_out=False):
        K, N = self.params.shape
        M = dy.size // N
        xt = dy_split = dyt = None
        if ops.bf16x3_ok(M, K, N):
            cached = getattr(self, "_xt", None)
            if cached is not None and cached[0] is x:
                xt = cached[1]
            if ops._mn_major():
                dh, dl, _ = ops.split_bf16(dy, M, N)               # one split of dy feeds dX and (MN-major) dW
                dy_split, dyt = (dh, dl), (dh, dl, 0)
            elif want_dx:
                dy_split, dyt = ops.split_bf16_both(dy, M, N)      # dy feeds dX as-is and dW transposed: one pass
        # dW / db feed nothing on the backward critical path: inside a captured step they run on a side branch
        lib.nt_side_begin()
        ops.linear_bwd_params(x, dy, self.params_delta, self.bias_delta, xt_split=xt, dyt_split=dyt)
        lib.nt_side_end()
        self._xt = None
        w_split, self._wsplit = getattr(self, "_wsplit", None), None
        if not want_dx:
            return None
        return ops.linear_bwd_input(dy, self.params, act, pre, addend, dy_split=dy_split, split_out=split_out, w_split=w_split)

    # --- reference protocol
    def forward(self, inputs):
        return self._forward(as_device(inputs, self.host_dtype))

    def backward(self, delta, inputs=[]):
        dy = as_device(delta, self.host_dtype)
        x = as_device(inputs, self.host_dtype) if len(inputs) else self.inputs
        return self._backward(dy, x)

    def setzero(self):
        self.params_delta.fill_zero()
        if self.bias:
            self.bias_delta.fill_zero()

    def update(self, lr=1e-10):
        self._wsplit = None                              # the cached bf16 split describes the weights before this update
        if self.adam:
            ops.adam_star(self.params, self.params_delta, self.moment_p, self.rmsprop_p, lr, self.t)
            if self.bias:
                ops.adam_star(self.bias_params, self.bias_delta, self.moment_b, self.rmsprop_b, lr, self.t)
            self.t += 1
        else:
            ops.sgd(self.params, self.params_delta, lr)
            if self.bias:
                ops.sgd(self.bias_params, self.bias_delta, lr)

    def _pad_out(self, mult=4):
        """Fast-path only (TrainStep, for the layer that feeds cross_entropy): pad the output width to a
        multiple of `mult` with dummy classes whose weights are 0 and whose bias is -1e30.  Their softmax
        probability is exactly 0, so loss, gradients, argmax and every real parameter are unchanged, while
        the padded pitch (e.g. vocab 2350 -> 2352, 10 -> 12) satisfies TMA's 16-byte stride rule and the
        three GEMMs of the layer run on the tensor cores instead of the SIMT fallback."""
        N = self.outfeature
        Np = (N + mult - 1) // mult * mult
        if Np == N or not self.bias:
            return
        w = np.zeros((self.infeature, Np), dtype=np.float32)
        w[:, :N] = self.params.numpy()
        b = np.full((Np,), -1e30, dtype=np.float32)
        b[:N] = self.bias_params.numpy()
        hd = self.host_dtype
        self.params = DeviceArray.from_host(w, host_dtype=hd)
        self.bias_params = DeviceArray.from_host(b, host_dtype=hd)
        self.params_delta = DeviceArray.zeros((self.infeature, Np), host_dtype=hd)
        self.bias_delta = DeviceArray.zeros((Np,), host_dtype=hd)
        if self.adam:
            self.moment_p, self.rmsprop_p = DeviceArray.zeros((self.infeature, Np)), DeviceArray.zeros((self.infeature, Np))
            self.moment_b, self.rmsprop_b = DeviceArray.zeros((Np,)), DeviceArray.zeros((Np,))
        self._logical_out, self.outfeature = N, Np
        from numpy_transformer_b200.device import bump_rehome
        bump_rehome()

    def save_model(self):
        n = getattr(self, "_logical_out", self.outfeature)
        if self.bias:
            return [np.asarray(self.params)[:, :n], np.asarray(self.bias_params)[:n]]
        return [np.asarray(self.params)[:, :n]]

    def _check_model(self, models):
        """Raises what the reference's `models[i].reshape(self.params.shape)` raises on a tree that is not this layer's
        (IndexError: too short, ValueError: wrong size) WITHOUT touching the parameters (fullconnect.py:161-164)."""
        n = getattr(self, "_logical_out", self.outfeature)
        np.asarray(models[0]).reshape(self.infeature, n)
        if self.bias:
            np.asarray(models[1]).reshape(n)

    def restore_model(self, models):
        self._check_model(models)
        n = getattr(self, "_logical_out", self.outfeature)
        w = np.asarray(models[0]).reshape(self.infeature, n)
        if n != self.outfeature:
            full = np.zeros(self.params.shape)
            full[:, :n] = w
            w = full
        self.params.set(w)
        self._wsplit = None
        if self.bias:
            b = np.asarray(models[1]).reshape(n)
            if n != self.outfeature:
                fb = np.full(self.bias_params.shape, -1e30)
                fb[:n] = b
                b = fb
            self.bias_params.set(b)

    def __name__(self):
        return "fclayer"

    def _slots(self):
        """(param, grad, m, v) attribute names for the trainer's flat arenas."""
        s = [("params", "params_delta", "moment_p", "rmsprop_p")]
        if self.bias:
            s.append(("bias_params", "bias_delta", "moment_b", "rmsprop_b"))
        return [(self, *x) for x in s]
