"""The tensor-core kernels must be Blackwell-native in the shipped binary: tcgen05.mma (SASS UTCHMMA) with TMEM loads
(LDTM) fed by TMA (UTMALDG / UTMASTG), and no legacy mma.sync (HMMA) anywhere.  Runs on the CPU box: cuobjdump only."""
import importlib.util
import os
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_gemm_kernels_are_tcgen05_tma_kernels(built_lib):
    spec = importlib.util.spec_from_file_location("sass_evidence", os.path.join(ROOT, "scripts", "sass_evidence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from selectivenet_for_semantic_segmentation_binary_b200 import _lib
    counts = mod.sass_counts(_lib.LIB_PATH)
    gemm = {n: c for n, c in counts.items()
            if any(k in n for k in ("conv3_halo2_kernel", "conv3_halo_kernel", "conv_gemm_kernel", "wgrad_gemm_kernel",
                                    "wgrad_gemm_pair_kernel", "wgrad64_kernel"))}
    assert len(gemm) >= 17                     # 6 + 3 + 5 + 3 instantiations
    for name, c in gemm.items():
        assert c["UTCHMMA"] > 0 and c["LDTM"] > 0 and c["UTMALDG"] > 0, name
        if "wgrad" not in name:
            assert c["UTMASTG"] > 0, name      # conv outputs leave through TMA stores
    assert all(c["HMMA"] == 0 for c in counts.values())
    # the CTA-pair kernels issue cta_group::2 MMAs
    import subprocess
    txt = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA.2CTA" in txt
