"""Host-side models of the prepared (not yet hardware-measured) render variants must keep matching the oracle: they are
what the kernels' index arithmetic and mask logic were written against (tools/check_tc2_layout.py, emulate_tc2.py,
emulate_tc3.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,expect", [
    ("check_tc2_layout.py", "tc3 operand layout ok"),
    ("emulate_tc2.py", "tc2 model matches the oracle"),
    ("emulate_tc3.py", "tc3 model matches the oracle"),
])
def test_model_script(script, expect):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert expect in r.stdout
