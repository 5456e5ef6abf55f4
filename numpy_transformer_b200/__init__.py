"""numpy_transformer_b200 — B200 (sm_100a) backend for the training step of ZouJiu1/numpy_transformer.

`install()` puts the drop-in mirror of the reference's modules (net.*, attention, classify,
PatchEmbed, Position_add, gpt.*) at the front of sys.path, so the reference's trainers import the
CUDA-backed classes with unchanged source.  There is no CPU fallback: device calls raise NtError
when libntb200.so is missing or no B200 is visible."""
import os
import sys

from ._lib import lib, NtError, LIBPATH  # noqa: F401
from .device import (DeviceArray, as_device, init, available, synchronize, launch_count,  # noqa: F401
                     set_gemm_mode, get_gemm_mode, PinnedBuffer)

DROPIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def install():
    """Make `import net.fullconnect`, `import attention`, `import attdecoderblock` ... resolve to the mirror."""
    for p in (os.path.join(DROPIN, "gpt"), DROPIN):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    stale = [m for m in sys.modules if m.split(".")[0] in
             ("net", "gpt", "attention", "classify", "PatchEmbed", "Position_add", "attdecoderblock")]
    for m in stale:
        f = getattr(sys.modules[m], "__file__", "") or ""
        if not f.startswith(DROPIN):
            del sys.modules[m]
    return DROPIN
