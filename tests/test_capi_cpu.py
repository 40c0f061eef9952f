"""No-GPU checks of the drop-in boundary: libmccba.so loads, exports every symbol include/mccba.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "mccba.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mccba_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    import multi_camera_calibration_b200 as m
    L = m.capi.lib()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), n
    assert set(m.capi.EXPORTS) <= set(names)


def test_defaults_follow_the_reference():
    import multi_camera_calibration_b200 as m
    L = m.capi.lib()
    o = m.capi.SolveOpts()
    assert L.mccba_default_solve_opts(C.byref(o)) == 0
    # TermCriteria(TermCriteria::COUNT, 20, 1e-7), include/opencv2/ccalib/multicalib.hpp:140
    assert (o.mode, o.crit_type, o.max_count) == (m.capi.MODE_REFERENCE_GN, m.capi.CRIT_COUNT, 20)
    assert o.epsilon == 1e-7
    opts = m.capi.Options()
    assert L.mccba_default_options(C.byref(opts)) == 0
    assert (opts.device, opts.rank, opts.nranks, opts.use_graph) == (0, 0, 1, 1)
    assert m.capi.PINHOLE == 0 and m.capi.OMNIDIRECTIONAL == 1      # multicalib.hpp:76-80


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import multi_camera_calibration_b200 as m
    with pytest.raises(m.MccbaError) as ei:
        m.Solver()
    assert ei.value.code == m.capi.ERR_CUDA


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "multi_camera_calibration_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or "import oracle" not in txt, f
                assert "from oracle" not in txt and "import oracle" not in txt and "mccba_oracle" not in txt, f
