"""A/B variants that have not run on a GPU yet.  They are compiled into the library behind
environment switches and leave the default kernels untouched (the SASS of the default
instantiations is byte-identical with and without them: `cuobjdump -sass` before / after,
recorded in the commit that adds each variant).  Their parity checks only run when
``IALS_EXPERIMENTAL=1`` so that an untested variant can never turn the round's suite red;
the first GPU call of the next round runs them (tools/gpu_round2_first.sh)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("IALS_EXPERIMENTAL") != "1",
                                 reason="unmeasured A/B variants: set IALS_EXPERIMENTAL=1")]


def _epochs(tmp_path, tag, extra_env):
    env = dict(os.environ)
    env.update(extra_env)
    out = str(tmp_path / f"{tag}.npz")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_epoch.py"), "--shape", "ml20m",
                    "--scale", "0.03", "--K", "128", "--epochs", "2", "--dump", out],
                   check=True, env=env, cwd=ROOT, timeout=600)
    return np.load(out)


def test_no_allocate_gather_is_bit_identical(tmp_path):
    """IALS_ROWS_LDG=na (cg_rows.cu ldg4_na): the same loads with L1::no_allocate -- only the
    cache policy differs, so two epochs must give the very same factors."""
    ref = _epochs(tmp_path, "default", {"IALS_ROWS_LDG": ""})
    na = _epochs(tmp_path, "na", {"IALS_ROWS_LDG": "na"})
    np.testing.assert_array_equal(ref["user"], na["user"])
    np.testing.assert_array_equal(ref["item"], na["item"])
