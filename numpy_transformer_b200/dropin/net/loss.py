"""net.loss on the B200 backend (mirrors net/loss.py:46-57, 68-71; the reference's binary_cross_entropy_loss :59-66
reads an undefined name and cannot run, so it has no counterpart)."""
import numpy as np
from numpy_transformer_b200._lib import lib
from numpy_transformer_b200.device import DeviceArray, as_device
from numpy_transformer_b200 import ops


def _labels_from_onehot(label):
    """Dense one-hot (rows, cols) ndarray -> int32 class indices, or None if it is not one-hot."""
    lab = np.asarray(label)
    if lab.ndim != 2:
        return None
    idx = lab.argmax(axis=-1)
    rows = np.arange(lab.shape[0])
    if not np.all(lab[rows, idx] == 1):
        return None
    if np.count_nonzero(lab) != lab.shape[0]:
        return None
    return idx.astype(np.int32)


def cross_entropy_loss(predict, label, n_norm=None):
    """-> (loss, partial, softmax) like net/loss.py:46-51.  `label` may be the reference's dense
    one-hot ndarray (converted to indices on the host), an int ndarray / int32 DeviceArray of class
    indices, or a dense float DeviceArray.  `loss` is a 0-d DeviceArray (float(loss) synchronises);
    np.argmax(softmax, -1) is answered from the device-side argmax without copying the matrix."""
    logits = as_device(predict)
    if logits.ndim != 2:
        raise ValueError("cross_entropy_loss expects (rows, classes), got %s" % (logits.shape,))
    labels = onehot = None
    if isinstance(label, DeviceArray):
        if label.dtype == np.int32:
            labels = label
        else:
            onehot = label
    else:
        lab = np.asarray(label)
        if lab.dtype.kind in "iu" and lab.ndim == 1:
            labels = DeviceArray.from_host(lab)
        else:
            idx = _labels_from_onehot(lab)
            if idx is not None:
                labels = DeviceArray.from_host(idx)
            else:
                onehot = DeviceArray.from_host(lab)
    loss, grad, prob, _ = ops.cross_entropy(logits, labels, onehot, n_norm)
    return loss, grad, prob


def mean_square_loss(predict, label):
    """net/loss.py:68-71: loss = mean((p - y)^2), partial = 2 (p - y) / size."""
    p = as_device(predict)
    l = as_device(label, p.host_dtype)
    assert p.size == l.size, "mean_square_loss: %s vs %s" % (p.shape, l.shape)
    loss, grad = DeviceArray((), host_dtype=np.float64), p.empty_like()
    lib.nt_mse_loss(p.ptr, l.ptr, loss.ptr, grad.ptr, p.size)
    return loss, grad


def binary_cross_entropy_logit_loss(predict, label):
    """net/loss.py:53-57: sigmoid inside the loss; returns (loss, partial, sigmoid)."""
    p = as_device(predict)
    l = as_device(label, p.host_dtype)
    assert p.size == l.size, "binary_cross_entropy_logit_loss: %s vs %s" % (p.shape, l.shape)
    loss, grad, sig = DeviceArray((), host_dtype=np.float64), p.empty_like(), p.empty_like()
    lib.nt_bce_logit_loss(p.ptr, l.ptr, loss.ptr, grad.ptr, sig.ptr, p.size)
    return loss, grad, sig
