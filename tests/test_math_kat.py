"""Known-answer sweep of the emulated libm (urban_road_filter_b200/csrc/urf_math.cuh, host build) against the container's
glibc: asinf/acosf over [-1,1], atanf over all bit patterns, atan2f on random pairs. Strided by default (seconds);
URF_FULL_KAT=1 runs every bit pattern (about a minute on 8 cores) — 0 mismatches in both."""
import os
import subprocess

from util import ROOT


def test_libm_emulation_sweep():
    full = os.environ.get("URF_FULL_KAT") == "1"
    stride, nrand = ("1", "1000000000") if full else ("509", "20000000")
    out = subprocess.run([os.path.join(ROOT, "build", "math_sweep"), stride, str(os.cpu_count() or 4), nrand],
                         capture_output=True, text=True, timeout=3600)
    print(out.stdout, out.stderr[-2000:])
    assert out.returncode == 0
    lines = dict(l.split(" ", 1) for l in out.stdout.strip().splitlines())
    for fn in ("asinf", "acosf", "atanf", "div_pi", "atan2f", "sector"):
        assert "mismatches=0" in lines[fn]
    checked, _, undecided = (int(v.split("=")[1]) for v in lines["sector"].split())
    assert undecided < 0.7 * checked          # most of the sample sits next to a boundary on purpose; uniform points: ~0.2 %


def test_std_sort_restatement_matches_libstdcxx():
    """urf_stdsort.cuh (what the star search's tie path runs on the device) against the real std::sort on the reference's
    record type and comparator: tie-heavy, sorted, reversed, organ-pipe, constant and median-of-three-killer arrays (the
    latter drive introsort into its heapsort fallback, which the binary reports having covered)."""
    out = subprocess.run([os.path.join(ROOT, "build", "stdsort_check"), "4000"], capture_output=True, text=True, timeout=900)
    print(out.stdout, out.stderr[-2000:])
    assert out.returncode == 0 and "mismatches=0" in out.stdout
    assert int(out.stdout.split("heapsort_fallbacks=")[1]) > 50
