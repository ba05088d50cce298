"""The near-first star sort (k_star_sort_warp / k_star_scan / k_star_refine) is exact: on random sectors, splitting at
the sampled pivot, walking the sorted near part and — without an edge there — resuming on the full order from the saved
running mean / deviation marks the same point as sorting everything and walking from the start (the reference's way,
star_shaped_search.cpp:109-150). Host check with the kernels' own arithmetic functions (tests/kat/star_prefix_check.cpp);
the kernels themselves are checked on the GPU by test_gpu_near_first_star_sort."""
import os
import re
import subprocess

import pytest

from util import ROOT


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_near_first_marks_the_same_point(seed):
    out = subprocess.run([os.path.join(ROOT, "build", "star_prefix_check"), "15000", str(seed)], capture_output=True, text=True, timeout=600)
    print(out.stdout, out.stderr[-2000:])
    assert out.returncode == 0
    f = {k: int(v) for k, v in re.findall(r"(\w+)=(\d+)", out.stdout)}
    assert f["mismatches"] == 0
    assert f["prefix_hits"] > 10000 and f["refined"] > 100          # both paths were exercised
