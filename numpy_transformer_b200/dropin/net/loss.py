"""net.loss on the B200 backend (mirrors net/loss.py:46-51, 66-71)."""
import numpy as np
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
    """net/loss.py:66-71 — host-side (not on the hot path of any BASELINE config)."""
    p, l = np.asarray(predict), np.asarray(label)
    loss = np.sum(np.square(p - l)) / p.size
    partial = 2 * (p - l) / p.size
    return loss, partial
