"""ctypes binding of libntb200.so.  Prototypes are parsed from include/ntb200.h so the header is
the single source of truth for the C ABI.  There is NO CPU fallback: if the library is missing
or no B200 is visible, importing works (so host-only logic is testable) but the first device
call raises."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(ROOT, "include", "ntb200.h")
LIBPATH = os.path.join(_HERE, "libntb200.so")

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "size_t": ctypes.c_size_t,
    "int32_t": ctypes.c_int32, "long long": ctypes.c_longlong,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every `int nt_*(...)` / `const char* nt_*(...)`."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int)\s+(nt_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    t = re.sub(r"\b(const|unsigned)\b", "", a).strip()
                    t = " ".join(t.split()[:-1])           # drop the parameter name
                    argtypes.append(_CTYPES[t])
        protos[name] = (ctypes.c_char_p if "char" in ret else ctypes.c_int, argtypes)
    return protos


class NtError(RuntimeError):
    pass


def _preload_nccl():
    """libntb200.so needs libnccl.so.2.  A process that later imports torch (the reference's transformer_of_image.py does,
    through torchvision) must end up with ONE NCCL: torch's libtorch_cuda.so is linked against the NCCL wheel bundled in
    site-packages (2.28) and fails with an undefined symbol if the older system library (2.27) was mapped first.  So map the
    bundled one first when it exists; libntb200 only uses the stable ncclCommInitRank / ncclAllReduce / ncclBroadcast ABI."""
    import glob
    import sys
    for base in sys.path:
        for cand in glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.*")):
            try:
                ctypes.CDLL(cand, mode=ctypes.RTLD_GLOBAL)
                return cand
            except OSError:
                pass
    return None


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()
        self._prof = None
        self._prof_external = False

    # ---- op-level CUDA-event profiler (bench.py roofline leg): every op call is bracketed by a
    # pair of events on the library stream; durations are read back in profile_end().
    _PROFILED = ("nt_linear_", "nt_layernorm_", "nt_attention_", "nt_cross_entropy", "nt_sgd", "nt_adam_star",
                 "nt_patchify", "nt_add", "nt_embedding_", "nt_gelu_", "nt_relu_", "nt_softmax_", "nt_dp_", "nt_vit_blocks_fwd",
                 "nt_vit_blocks_bwd", "nt_split_bf16", "nt_kv_append", "nt_decode_embed",
                 "nt_argmax_next", "nt_increment", "nt_p2p_allreduce", "nt_vit_tc_fwd", "nt_vit_tc_bwd")

    def profile_begin(self, external=False):
        """external=True records with cudaEventRecordExternal so the pairs survive stream capture as graph
        nodes (per-op device times of a replayed graph, no host gaps between ops)."""
        self._prof = []
        self._prof_external = external

    def profile_pause(self):
        recs, self._prof = self._prof, None
        return recs

    def profile_read(self, recs):
        import ctypes as C
        dll = self.load()
        out = []
        for name, args, e0, e1 in recs:
            ms = C.c_float(0)
            dll.nt_event_elapsed_ms(e0, e1, C.byref(ms))
            out.append((name, args, ms.value))
        return out

    def profile_end(self):
        import ctypes as C
        recs, self._prof = self._prof, None
        dll = self.load()
        out = []
        for name, args, e0, e1 in recs:
            ms = C.c_float(0)
            dll.nt_event_elapsed_ms(e0, e1, C.byref(ms))
            dll.nt_event_destroy(e0)
            dll.nt_event_destroy(e1)
            out.append((name, args, ms.value))
        return out

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIBPATH):
            raise NtError("libntb200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C numpy_transformer_b200/csrc`. This backend has no CPU fallback." % LIBPATH)
        _preload_nccl()
        dll = ctypes.CDLL(LIBPATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in self.protos.items():
            fn = getattr(dll, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        self._dll = dll
        return dll

    def __getattr__(self, name):
        if name.startswith("nt_"):
            fn = getattr(self.load(), name)
            if fn.restype is ctypes.c_char_p:
                return fn

            profiled = name.startswith(self._PROFILED)

            def call(*a):
                if self._prof is not None and profiled:
                    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
                    self._dll.nt_event_create(ctypes.byref(e0))
                    self._dll.nt_event_create(ctypes.byref(e1))
                    rec = self._dll.nt_event_record_external if self._prof_external else self._dll.nt_event_record
                    rec(e0)
                    rc = fn(*a)
                    rec(e1)
                    self._prof.append((name, tuple(x for x in a if isinstance(x, float) or (isinstance(x, int) and x < (1 << 31))), e0, e1))
                else:
                    rc = fn(*a)
                if rc != 0:
                    raise NtError("%s failed (%d): %s" % (name, rc, (self._dll.nt_last_error() or b"").decode()))
            call.__name__ = name
            setattr(self, name, call)
            return call
        raise AttributeError(name)


lib = _Lib()
