"""GPU: back-to-back launches of the pipelined kernels at assorted shapes without host synchronisation (protocol races
such as mbarrier phase aliasing only show up under such load); results must not drift."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_kernel_soak():
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "soak.py"), "25"], capture_output=True, text=True, cwd=ROOT,
                       timeout=600)
    assert r.returncode == 0 and "SOAK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
