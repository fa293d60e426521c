"""Randomised parity of the sliding-window kernel against the oracle (scripts/stress_dense.py): random sizes incl. widths that
are not multiples of 4, all template widths the kernel covers, heights 1..64 (both ring layouts), both camera sides, negative /
empty / unbounded disparity ranges, flat frames (all ties), random accept thresholds. Bit-exact, every case."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [7, 8])
def test_dense_kernel_random_configs(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "stress_dense.py"), "90", str(seed)], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatches 0" in r.stdout


@pytest.mark.parametrize("seed", [11])
def test_dense_correlation_kernel_random_configs(seed):
    """NCC / ZNCC / SSD (gray and colour) and colour SAD on the sliding-window correlation kernel: indices, costs, f64 scores and
    MatchValues bit-exact."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "stress_corr.py"), "70", str(seed)], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatches 0" in r.stdout and "kernels 70," in r.stdout
