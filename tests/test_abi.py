"""CPU: the C-ABI library loads and exports exactly what include/mnemo_cuda.h declares; no compute."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mnemo_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mnemophonix_b200 import capi
    L = capi.lib()
    names = declared_symbols(os.path.join(ROOT, "include", "mnemo_cuda.h"))
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_size_functions_follow_reference_formulas():
    from mnemophonix_b200 import capi
    L = capi.lib()
    assert L.mnemo_samples_5512(1323000) == 165375
    assert L.mnemo_frames(165375) == 2552 and L.mnemo_images(165375) == 304
    assert L.mnemo_frames(31444) == 460 and L.mnemo_images(31444) == 42            # README.md:52-56
    assert L.mnemo_frames(2047) == 0 and L.mnemo_images(10175) == 0                # FILE_TOO_SMALL cases
    assert L.mnemo_images(10176) == 1
    assert L.mnemo_images(39690000) == 77500                                       # 2 h (C2)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    from mnemophonix_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.MnemoError):
        capi.Context(0)


def test_product_never_touches_the_oracle():
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "mnemophonix_b200")):
        for f in fs:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"\boracle\b", txt) and "mnemo_oracle" in txt or re.search(r"(from|import)\s+oracle", txt):
                    bad.append(f)
    assert not bad, bad


def test_cuda_library_is_sm100a_only():
    so = os.path.join(ROOT, "mnemophonix_b200", "libmnemo_cuda.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
