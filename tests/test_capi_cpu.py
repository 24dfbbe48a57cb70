"""CPU tests of the product's native boundary: the library loads, exports every symbol the headers
declare, and refuses to run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_functions(header):
    txt = (ROOT / "include" / header).read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(msckf_(?:b200|mono)_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("header", ["msckf_b200.h", "msckf_mono_c.h"])
def test_library_exports_every_declared_symbol(engine_lib, header):
    lib = C.CDLL(str(engine_lib))
    names = declared_functions(header)
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/{header} but not exported"


def test_no_cpu_fallback_without_device(engine_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from msckf_mono_b200 import engine_filter, synth
    f = engine_filter(np.float32)
    wl = synth.make_window_workload(n_features=3, n_clones=4, seq=5)
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback|CUDA"):
        f.initialize(wl["camera"], wl["noise"], wl["params"], wl["imu_state"])


def test_product_sources_never_reference_the_oracle():
    for p in list((ROOT / "msckf_mono_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".cpp", ".h", ".hpp"}:
            txt = p.read_text()
            assert "msckf_oracle" not in txt.replace("msckf_oracle_", "X") or p.name == "cview.py", p
            assert "oracle/" not in txt or p.name in {"cview.py", "__init__.py"} or "gen_chi2" in p.name, p
