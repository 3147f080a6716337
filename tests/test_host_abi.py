"""CPU-side checks of the product: the C-ABI library loads and exports every symbol that
include/orb_b200.h declares, the host-side plan reproduces the reference's geometry, and the
product fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orb_slam2_chinesenotes_b200 as ob
from oracle_lib import OracleExtractor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "orb_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(orb[xm]_[a-z0-9_]+)\s*\(", text)))


def test_abi_exports_every_declared_symbol(built_lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(built_lib, s), f"{s} declared in include/orb_b200.h but not exported"


def test_product_does_not_link_the_oracle(built_lib):
    import subprocess
    out = subprocess.run(["ldd", ob.LIB_PATH], capture_output=True, text=True).stdout
    assert "orb_oracle" not in out and "orbref" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "orb_slam2_chinesenotes_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")):
                src = open(os.path.join(root, f)).read()
                assert "oracle/" not in src.replace("no CPU fallback", "") or f == "__init__.py" and "oracle" not in src, f


@pytest.mark.parametrize("w,h,nf,cells,quota", [
    (640, 480, 1000, [280, 192, 130, 88, 54, 35, 24, 12], [217, 181, 151, 126, 105, 87, 73, 60]),
    (1241, 376, 2000, [429, 297, 189, 132, 72, 45, 36, 20], [434, 362, 302, 251, 209, 175, 145, 122]),
    (752, 480, 1200, [336, 228, 160, 104, 66, 45, 28, 15], [261, 217, 181, 151, 126, 105, 87, 72]),
    (1920, 1080, 4000, [2074, 1428, 966, 665, 464, 312, 220, 128], [869, 724, 603, 503, 419, 349, 291, 242]),
])
def test_plan_matches_reference_geometry(built_lib, w, h, nf, cells, quota):
    """level sizes, processed 30-px cells, per-level quotas."""
    d = ob.plan_describe(nf, 1.2, 8, 20, 7, w, h)
    assert d["cells"].tolist() == cells and d["quota"].tolist() == quota
    from synth import synth_frame
    O = OracleExtractor(nf)
    O.extract(synth_frame(w, h, 1)) if w * h < 700000 else None
    if w * h < 700000:
        for l in range(8):
            assert O.level_size(l) == (int(d["level_w"][l]), int(d["level_h"][l]))
        # the provable candidate bound holds
        for l in range(8):
            assert len(O.candidates(l)) <= d["cand_cap"][l]
    O.close()


def test_unsupported_shapes_are_rejected(built_lib):
    for (w, h) in [(60, 60), (100, 400), (5000, 400)]:
        with pytest.raises(ob.OrbError):
            ob.plan_describe(1000, 1.2, 8, 20, 7, w, h)


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ob.OrbError) as e:
        ob.ORBextractor(1000, 1.2, 8, 20, 7)
    assert e.value.code == ob.ORBX_E_CUDA
    with pytest.raises(ob.OrbError):
        ob.hamming_bf(np.zeros((4, 32), np.uint8), np.zeros((4, 32), np.uint8))


def test_max_keypoints_bounds_the_reference_counts(built_lib):
    """orbx_max_keypoints (host only) is an upper bound on what operator() returns: checked on the committed outputs of
    the reference's own extractor."""
    import os
    import numpy as np
    import orb_slam2_chinesenotes_b200 as ob
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in ("tum", "kitti", "small"):
        g = np.load(os.path.join(gdir, f"ref_extract_{name}.npz"))
        bound = ob.max_keypoints(int(g["nfeatures"]), 1.2, 8, 20, 7, int(g["w"]), int(g["h"]))
        assert len(g["kps"]) <= bound <= int(g["nfeatures"]) + 4 * 8 + 64, (name, len(g["kps"]), bound)
    with pytest.raises(ob.OrbError):
        ob.max_keypoints(1000, 1.2, 8, 20, 7, 20, 20)
