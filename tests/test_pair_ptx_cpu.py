"""Groundwork for the CTA-pair convolution kernel: the cta_group::2 / multicast PTX wrappers in
csrc/experimental/seg_ptx_pair.cuh must assemble for sm_100a and produce the 2-CTA SASS forms.  Compile-only (no GPU)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not (os.path.exists("/usr/local/cuda/bin/nvcc") or shutil.which("nvcc")), reason="nvcc not available")
def test_pair_wrappers_assemble_for_sm100a(tmp_path):
    env = dict(os.environ, TMPDIR=str(tmp_path))
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "check_pair_ptx.sh")], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTMALDG.4D.IM2COL.2CTA", "UTMALDG.2D.MULTICAST", "UTCBAR.2CTA.MULTICAST", "UCGABAR_ARV"):
        assert mnemonic in r.stdout, (mnemonic, r.stdout[-1500:])
