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


@pytest.mark.parametrize("seed", [21])
def test_tensor_pipe_correlation_kernel_random_configs(seed):
    """NCC / ZNCC / SSD on the tensor-pipe kernel (usv_dense_mma.cu: IMMA row products, sliding accumulators): template widths
    1..32 incl. odd ones, any height, gray and colour, frames wide enough for several passes over the candidate columns, both
    camera sides, bounded / negative / unbounded ranges, flat frames. Bit-exact, every case, and every case on that kernel."""
    # forced: with the automatic dispatch one-plane NCC / ZNCC on wide frames would go to the tcgen05 kernel
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "stress_corr.py"), "70", str(seed), "mma", "mma"], cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatches 0" in r.stdout and "'dense_corr_mma_kernel': 70" in r.stdout


def test_alu_correlation_kernel_still_matches():
    """usv_set_option(USV_OPT_CORR_KERNEL, ALU) keeps the correlation sweeps on the ALU kernel (the A/B arm of the measurements): still bit-exact."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "stress_corr.py"), "40", "5", "all", "alu"], cwd=ROOT, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatches 0" in r.stdout and "dense_corr_mma_kernel" not in r.stdout


def test_tcgen05_correlation_kernel_random_configs():
    """USV_OPT_CORR_KERNEL = TCGEN05 routes the correlation sweeps to the tcgen05 kernel (usv_dense_umma.cu: UMMA kind::i8 products,
    accumulators in tensor memory, operand tiles in the canonical K-major layout) wherever it applies (unset, only where it
    is measured faster: one-plane NCC / ZNCC on wide frames). Every case bit-exact, every case on that kernel."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "stress_corr.py"), "60", "33", "mma", "tcgen05"], cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatches 0" in r.stdout and "'dense_corr_umma_kernel': 60" in r.stdout


def test_default_dispatch_mixes_the_tensor_kernels():
    """With the automatic dispatch the correlation sweeps split between the mma.sync kernel (colour, SSD, narrow frames) and the
    tcgen05 kernel (one-plane NCC / ZNCC on frames at least 128 windows wide): both appear, nothing mismatches."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "stress_corr.py"), "120", "44", "mma"], cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatches 0" in r.stdout and "dense_corr_mma_kernel" in r.stdout and "dense_corr_umma_kernel" in r.stdout
