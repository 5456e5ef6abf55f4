"""Batch construction on the device.

Once the training step takes half a millisecond, building the batch on the host — the fancy-index shuffle and slice of
transformer_of_image.py:128-137, the list comprehensions of getinputs (gpt/gpt_train_potry3000.py:106-133), the dense
one-hot labels — costs more than the step.  Here the data set is uploaded ONCE (MNIST as uint8: 47 MB; a corpus as
int32 token ids) and a step's batch is one gather kernel writing straight into `TrainStep.d_in / d_lab`; the only
per-step host traffic is the random draw itself (a permutation per epoch, `batch` offsets per GPT step), taken from
the same `np.random` calls as the reference so a seeded run sees the same batches."""
import numpy as np

from ._lib import lib
from .device import DeviceArray, _Buffer


class ImageBatches:
    """The epoch loop of transformer_of_image.py:97-138 for uint8 images [N, H, W] (or [N, C, H, W]) and int labels [N]:
    `/255` normalisation (:99-100), `np.random.shuffle` of the index vector once per epoch (:128-133), batch j = rows
    j*B .. (j+1)*B of the shuffled set (:135-139).  `batch(j, images_out, labels_out)` fills device arrays."""

    def __init__(self, images_u8, labels, batch):
        img = np.ascontiguousarray(images_u8)
        assert img.dtype == np.uint8, "ImageBatches keeps the raw uint8 pixels; normalisation happens in the gather"
        self.n = img.shape[0]
        self.sample_shape = img.shape[1:] if img.ndim == 4 else (1,) + img.shape[1:]
        self.pixels = int(np.prod(self.sample_shape))
        self.batch = int(batch)
        self._pix = _Buffer(img.nbytes)
        lib.nt_h2d(self._pix.ptr, img.ctypes.data, img.nbytes)
        self.labels = DeviceArray.from_host(np.asarray(labels).astype(np.int32))
        # the reference divides in fp64 and the upload casts to fp32: one rounding chain per pixel value, tabulated
        self.lut = DeviceArray.from_host((np.arange(256, dtype=np.float64) / 255).astype(np.float32))
        self.order = np.arange(self.n)
        self.perm = DeviceArray.from_host(self.order.astype(np.int32))

    def __len__(self):
        return self.n // self.batch                     # full batches per epoch

    def shuffle(self):
        """`k = np.arange(n); np.random.shuffle(k); datas = datas[k]` (:128-133) — the shuffles compose across epochs."""
        k = np.arange(self.n)
        np.random.shuffle(k)
        self.order = self.order[k]
        self.perm.set(self.order.astype(np.int32))

    def batch_into(self, j, images_out, labels_out):
        first = j * self.batch
        assert 0 <= first and first + self.batch <= self.n, "batch %d out of range" % j
        assert images_out.size == self.batch * self.pixels and labels_out.size == self.batch
        lib.nt_gather_images_u8(self._pix.ptr, self.perm.ptr, first, self.lut.ptr, images_out.ptr, self.batch, self.pixels)
        lib.nt_gather_i32(self.labels.ptr, self.perm.ptr, first, labels_out.ptr, self.batch)

    def feed(self, j, train_step):
        """Batch j straight into a TrainStep's device input buffers; follow with `train_step.run_device()`."""
        self.batch_into(j, train_step.d_in, train_step.d_lab)


class TokenWindows:
    """getinputs (gpt/gpt_train_potry3000.py:106-133) for a corpus already mapped to token ids: `id_start =
    np.random.choice(inputid, batch, replace=False)`, inputs = ids[s : s+T], labels = ids[s+1 : s+T+1]."""

    def __init__(self, token_ids, context, batch):
        ids = np.asarray(token_ids).astype(np.int32).reshape(-1)
        self.n, self.context, self.batch = ids.size, int(context), int(batch)
        self.corpus = DeviceArray.from_host(ids)
        self.start = DeviceArray.zeros((self.batch,), np.int32)

    def sample_into(self, inputid, inputs_out, labels_out):
        id_start = np.random.choice(inputid, self.batch, replace=False)
        assert id_start.max() + self.context + 1 <= self.n, "window runs past the corpus"
        self.start.set(id_start.astype(np.int32))
        assert inputs_out.size == self.batch * self.context and labels_out.size == self.batch * self.context
        lib.nt_gather_windows(self.corpus.ptr, self.start.ptr, inputs_out.ptr, labels_out.ptr, self.batch, self.context)
        return id_start

    def feed(self, inputid, train_step):
        return self.sample_into(inputid, train_step.d_in, train_step.d_lab)
