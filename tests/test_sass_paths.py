"""CPU: the shipped extension really contains the hardware paths DESIGN.md claims (cuobjdump -sass of the built .so):
fp64 tensor-core DMMA in every blocked factorisation kernel, legacy-tensor HMMA (TF32) in the k-NN prefilter, cp.async
(LDGSTS) operand staging, and nothing from tcgen05 / TMA (fp64 has no tcgen05 kind)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "cgp_b200", "_cgp_b200.so")


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(SO), reason="needs cuobjdump and the built .so")
def test_tensor_core_paths_in_sass():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True, timeout=300).stdout
    assert "sm_100a" in out
    per = {}
    cur = None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = {"DMMA": 0, "HMMA": 0, "LDGSTS": 0, "UTMALDG": 0, "UTC": 0}
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            for k in per[cur]:
                if m.group(1).startswith(k):
                    per[cur][k] += 1

    def total(substr, key):
        return sum(v[key] for n, v in per.items() if substr in n)

    for kern in ("k_lauum", "k_potrf_trail", "k_potrf_head", "k_potrf_update", "k_potrf_panel", "k_trtri_trail", "k_trtri_acc",
                 "k_trtri_scale", "k_predict_vnorm", "k_dgemm_nt"):
        assert total(kern, "DMMA") >= 256, kern   # the unrolled 128x128 (or 64x64) DMMA mainloop
        assert total(kern, "LDGSTS") > 0, kern     # cp.async operand ring
    assert total("k_knn_tc", "HMMA") >= 6          # 3 TF32 MMAs per tile, two template instances
    assert total("k_knn_tc", "DMMA") == 0
    assert sum(v["UTMALDG"] + v["UTC"] for v in per.values()) == 0
    assert sum(v["DMMA"] for v in per.values()) >= 4000
