"""CPU-only checks: the C-ABI library loads and exports every symbol include/ntb200.h declares,
the product path fails loudly without a GPU (no CPU fallback), and host-side helpers."""
import os
import subprocess
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nt():
    import numpy_transformer_b200 as nt
    if not os.path.exists(nt.LIBPATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "numpy_transformer_b200", "csrc"), "-j8"])
    return nt


def test_header_parses_and_library_exports_every_symbol(nt):
    protos = nt.lib.protos
    assert len(protos) >= 50
    for must in ("nt_linear_fwd", "nt_linear_bwd_input", "nt_linear_bwd_params", "nt_layernorm_fwd",
                 "nt_attention_fwd", "nt_attention_bwd", "nt_cross_entropy", "nt_adam_star", "nt_dp_allreduce_sum"):
        assert must in protos
    dll = nt.lib.load()
    for name in protos:
        assert hasattr(dll, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", nt.LIBPATH]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(protos) <= exported


def test_library_is_sm100a_only_and_carries_tcgen05_tma(nt):
    elf = subprocess.run(["cuobjdump", "-lelf", nt.LIBPATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    assert "sm_90" not in elf and "sm_80" not in elf
    # the Blackwell-native instructions themselves, per kernel (SASS mnemonics: tcgen05.mma -> UTC*MMA, tcgen05.ld ->
    # LDTM, TMA tensor loads -> UTMALDG, bulk copies -> UBLKCP; /opt/skills/guides/B200_PROFILING.md)
    sass = subprocess.run(["cuobjdump", "-sass", nt.LIBPATH], capture_output=True, text=True).stdout
    per_kernel, cur = {}, None
    for line in sass.splitlines():
        line = line.strip()
        if line.startswith("Function :"):
            cur = per_kernel.setdefault(line.split(":", 1)[1].strip(), {"UTCHMMA": 0, "LDTM": 0, "UTMALDG": 0, "UBLKCP": 0, "HMMA": 0,
                                                                        "UCGABAR_ARV": 0, "UCGABAR_WAIT": 0})
        elif cur is not None:
            for op in cur:
                if " " + op in line or line.startswith(op):
                    if op == "HMMA" and "UTCHMMA" in line:
                        continue
                    cur[op] += 1

    def total(pattern, op):
        return sum(v[op] for k, v in per_kernel.items() if pattern in k)
    assert total("gemm_tc_kernel", "UTCHMMA") > 0 and total("gemm_tc_kernel", "LDTM") > 0 and total("gemm_tc_kernel", "UTMALDG") > 0
    # the fused ViT stack (the headline kernel) and the dh=64 attention kernels run their contractions on tcgen05
    for k in ("vit_tc_fwd", "vit_tc_bwd", "attn_tc_kernel", "attn_tcl_fwd_kernel"):
        if any(k in name for name in per_kernel):
            assert total(k, "UTCHMMA") > 0 and total(k, "LDTM") > 0 and total(k, "HMMA") == 0, k
    # the default attention forward of the GPT shapes takes its operands by TMA
    assert total("attn_tcl_fwd_kernel", "UTMALDG") > 0
    # the phase-program kernel synchronises its CTAs with the hardware cluster barrier (barrier.cluster.arrive / wait)
    assert total("phase_program_kernel", "UCGABAR_ARV") > 0 and total("phase_program_kernel", "UCGABAR_WAIT") > 0


def test_reference_tree_is_the_unmodified_reference():
    """oracle/_ref/tree (what the GPU tests execute and bench.py --impl reference times) is byte-identical to the
    reference: every file's sha256 equals tests/golden/ref_scripts/MANIFEST.json, which was written from /root/reference."""
    import json
    from oracle import ref_tree
    want = json.load(open(ref_tree.MANIFEST))
    assert "transformer_of_image.py" in want and "gpt/gpt_train_potry3000.py" in want and "net/activation.py" in want
    if os.path.isdir(ref_tree.REFERENCE):
        for rel, sha in want.items():
            assert ref_tree._sha(os.path.join(ref_tree.REFERENCE, rel)) == sha, rel
    if not ref_tree.available():
        pytest.skip("oracle/_ref/tree has not been built here (python -c 'import __graft_entry__ as g; g.build()')")
    assert ref_tree.tree_manifest() == want


def test_golden_trainer_logs_parse():
    from oracle import ref_tree
    gold = os.path.dirname(ref_tree.MANIFEST)
    rows, extra = ref_tree.parse_log(open(os.path.join(gold, "vit.log"), encoding="utf-8").read())
    assert len(rows) == 22 and rows[0]["epoch"] == 25 and rows[-1]["epoch"] == 35
    assert sum("testset precision" in l for l in extra) == 11
    rows, _ = ref_tree.parse_log(open(os.path.join(gold, "gpt_char.log"), encoding="utf-8").read())
    assert len(rows) == 101 and rows[0]["lr"] == 0.0 and rows[-1]["loss"] < 0.5 * rows[0]["loss"]
    rows, _ = ref_tree.parse_log(open(os.path.join(gold, "gpt_poetry.log"), encoding="utf-8").read())
    assert [r["iters"] for r in rows] == ["6599_6600", "6600_6600", "6601_6600"] and rows[1]["valpre"] is not None


def test_no_cpu_fallback(nt):
    if nt.available():
        pytest.skip("a GPU is visible")
    nt.install()
    from net.fullconnect import fclayer
    with pytest.raises(nt.NtError):
        fclayer(3, 4, True)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "numpy_transformer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src, os.path.join(dp, f)


def test_mask_detection_and_onehot_conversion(nt):
    from numpy_transformer_b200 import ops
    nt.install()
    from net.loss import _labels_from_onehot
    assert ops.detect_mask([], 4) == (ops.MASK_NONE, None)
    m = np.tril(np.ones((4, 4)))
    c = np.where(m == 0, -np.inf, 0.0)
    assert ops.detect_mask(c, 4) == (ops.MASK_CAUSAL, None)
    assert ops.detect_mask(np.zeros((4, 4)), 4) == (ops.MASK_NONE, None)
    assert ops.detect_mask("causal", 4) == (ops.MASK_CAUSAL, None)
    oh = np.zeros((5, 7))
    oh[np.arange(5), [1, 0, 6, 3, 3]] = 1
    assert _labels_from_onehot(oh).tolist() == [1, 0, 6, 3, 3]
    assert _labels_from_onehot(oh * 0.5) is None
    oh[0, 2] = 1
    assert _labels_from_onehot(oh) is None


def test_dropin_module_surface(nt):
    """Every import the reference's trainers perform resolves to the mirror."""
    nt.install()
    import importlib
    for mod, names in [("net.fullconnect", ["fclayer"]), ("net.layernorm", ["layer_norm"]),
                       ("net.activation", ["ReLU", "sigmoid", "tanh", "SiLU", "Softmax", "GELU"]),
                       ("net.loss", ["cross_entropy_loss", "binary_cross_entropy_logit_loss", "mean_square_loss"]),
                       ("net.flatten", ["flatten_layer"]), ("net.Embedding", ["Embedding_layer"]),
                       ("net.Convolution", ["convolution_layer"]), ("attention", ["attention_layer"]),
                       ("classify", ["classify_layer"]),
                       ("PatchEmbed", ["PatchEmbed_flatten", "PatchEmbed_convolution", "Position_Embedding"]),
                       ("Position_add", ["Position_learnable"]), ("attdecoderblock", ["attdecoderblock_layer"]),
                       ("gpt.attdecoderblock", ["attdecoderblock_layer"]), ("gpt.classify_gpt", ["classify_layer"]),
                       ("gpt.gpt_linear", ["gpt_linear_layer"]), ("gpt.FixedEmbed", ["Position_Fixed"])]:
        m = importlib.import_module(mod)
        assert m.__file__.startswith(nt.DROPIN), mod
        for n in names:
            assert hasattr(m, n), (mod, n)
    from net.flatten import flatten_layer
    from net.activation import ReLU
    for cls in (flatten_layer, ReLU):
        assert "save_model" not in dir(cls) and "restore_model" not in dir(cls)
    from gpt.attdecoderblock import attdecoderblock_layer
    assert "attdecoderblock" in attdecoderblock_layer.__module__          # gpt_train_potry3000.py:220-221


def test_sincos_table_matches_oracle(nt):
    nt.install()
    from gpt.FixedEmbed import sincos_table
    from oracle.numpy_ref import posk_table
    assert np.array_equal(sincos_table(49, 96)[None], posk_table(49, 96))


def test_python_block_cache_uses_the_pool_size_classes():
    """device._size_class must mirror csrc/runtime.cu:size_class — blocks recycled on the Python side are handed out by
    class, so a mismatch would hand a too-small block to a larger tensor."""
    import re
    from numpy_transformer_b200.device import _size_class
    src = open(os.path.join(ROOT, "numpy_transformer_b200", "csrc", "runtime.cu")).read()
    body = src[src.index("static size_t size_class(size_t b)"):]
    body = body[:body.index("}") + 1]
    assert re.search(r"b < 512\) return 512", body) and "(1u << 20)" in body and "(1u << 16)" in body and "511" in body

    def c_size_class(b):                       # the C function, restated from the text checked above
        if b < 512:
            return 512
        if b < (1 << 20):
            return (b + 511) & ~511
        return (b + (1 << 16) - 1) & ~((1 << 16) - 1)
    for b in [1, 4, 511, 512, 513, 4096, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, 77 * 1024 * 1024 + 3]:
        assert _size_class(b) == c_size_class(b) >= b
