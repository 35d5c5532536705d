"""The C-ABI shared library loads here (no GPU) and exports every symbol include/sunet_b200.h declares;
the ctypes signature table covers exactly that set.  No compute calls."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sunet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sunet_[a-z0-9_A-Z]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    names = _declared()
    assert len(names) >= 24
    raw = ctypes.CDLL(os.path.join(ROOT, "selectivenet_for_semantic_segmentation_binary_b200", "libsunet_b200.so"))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"


def test_binding_table_matches_header(built_lib):
    from selectivenet_for_semantic_segmentation_binary_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    assert built_lib.sunet_abi_version() == 2
    assert built_lib.sunet_launch_count() == 0


def test_argument_validation_without_gpu(built_lib):
    """Bad arguments are rejected on the host before anything touches CUDA, with an error string."""
    from selectivenet_for_semantic_segmentation_binary_b200 import _lib
    a = _lib.ConvGemmArgs()
    assert built_lib.sunet_conv_gemm(ctypes.byref(a), None) != 0
    assert b"conv_gemm" in built_lib.sunet_last_error()
    assert built_lib.sunet_conv_gemm_stat_rows(ctypes.byref(a)) == -1
    assert built_lib.sunet_metric_hist(None, None, None, 0, 0, 0.0, 0.0, 0, None, None) != 0
    assert built_lib.sunet_adam_step(None, 0, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, None, None, None) != 0


def test_missing_library_fails_loudly(monkeypatch):
    from selectivenet_for_semantic_segmentation_binary_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsunet_b200.so")
    import pytest
    with pytest.raises(_lib.SunetError):
        _lib.load()


def _header_struct_fields(name):
    src = open(os.path.join(ROOT, "include", "sunet_b200.h")).read()
    body = src[src.index(f"typedef struct {name} {{"):src.index(f"}} {name};")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    body = body[body.index("{") + 1:]
    return re.findall(r"(\w+)\s*(?:,|;)", body)


def test_ctypes_structs_mirror_the_header_field_for_field():
    """A field added to a C struct without its ctypes twin would shift every later argument silently."""
    from selectivenet_for_semantic_segmentation_binary_b200 import _lib
    pairs = [("sunet_conv_gemm_args", _lib.ConvGemmArgs), ("sunet_wgrad_gemm_args", _lib.WgradGemmArgs),
             ("sunet_pack_job", _lib.PackJob), ("sunet_adam_tensor", _lib.AdamTensor)]
    for cname, cls in pairs:
        assert _header_struct_fields(cname) == [f[0] for f in cls._fields_], cname
