"""Compile-time development variants of libalphagpu.so (alphagpu_b200/libalphagpu_<name>.so, built by
scripts/variants_build_and_compare.sh) must reproduce the default library's self-play output bit for bit.  Skipped when no variant
library is present in the tree (the usual state: variants are built on demand and are not part of the product)."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = sorted(glob.glob(os.path.join(ROOT, "alphagpu_b200", "libalphagpu_*.so")))

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not VARIANTS, reason="no variant libraries built")
@pytest.mark.parametrize("lib", VARIANTS or ["none"], ids=lambda p: os.path.basename(p))
def test_variant_library_reproduces_the_default_output(lib):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "lib_variant_experiment.py"), lib, "--games", "8192", "--reps", "1"],
                         capture_output=True, text=True, timeout=600)
    last = json.loads(out.stdout.strip().splitlines()[-1])
    assert out.returncode == 0 and last == {"all_identical": True}, out.stdout[-2000:] + out.stderr[-2000:]
