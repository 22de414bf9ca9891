"""The exact-arithmetic contract at the instruction level (CPU: cuobjdump on the built objects).

K1 replays the reference's FFT butterflies with every multiply and add rounded on its own (fft.c:53-62).  ptxas contracts
a packed multiply feeding a packed add into FFMA2 even with -fmad=false, so the kernel never writes that pattern; the
only fused instructions it may contain are the FMA-by-a-run-time-one of twiddle_mul (k1_spectral.cu), one per complex
product, i.e. one FFMA2 for every two FMUL2 of a butterfly.  A scalar FFMA, or more FFMA2 than that, means the compiler
fused something the reference rounds separately."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def hist():
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    import sass_histogram
    h = sass_histogram.histograms()
    assert h, "no objects under mnemophonix_b200/_obj: run python -m mnemophonix_b200.build"
    return h


def test_k1_has_no_unintended_fused_multiply_add(hist):
    k1 = hist["k1_spectral_kernel"]
    # the only scalar FFMAs are the Newton steps of the one IEEE division per frame (logbins.c:75, __fdiv_rn: 5 on its
    # fast path, 12 in its out-of-range subroutine); the ~900 unrolled butterflies would show up as hundreds.  (HFMA2 with
    # zero operands is ptxas' idiom for loading a zero.)
    assert k1.get("FFMA", 0) <= 24, "scalar FFMA in K1: a product was fused with an add the reference rounds separately"
    assert k1.get("DFMA", 0) == 0
    assert k1["FMUL2"] > 500 and k1["FADD2"] > 500, "K1 lost its packed f32x2 butterflies"
    # twiddle_mul = 2 FMUL2 + 1 FFMA2; the other FMUL2 are the 1/1024 scalings and squares of stage 11
    assert 0 < k1["FFMA2"] <= k1["FMUL2"] // 2, (k1["FFMA2"], k1["FMUL2"])


def test_k1_stages_tiles_with_tma_bulk_copies(hist):
    k1 = hist["k1_spectral_kernel"]
    assert k1.get("UBLKCP", 0) >= 1 and k1.get("SYNCS", 0) >= 1      # cp.async.bulk + mbarrier


def test_no_tensor_core_or_library_code_on_the_path(hist):
    for name, c in hist.items():
        assert not any(op.startswith(("HMMA", "IMMA", "UTCMMA", "UTCHMMA", "QGMMA")) for op in c), name
