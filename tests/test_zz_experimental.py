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


def test_kmajor_gram_passes_the_gram_and_heavy_row_parity_tests():
    """IALS_WGRAM=kmajor (wgram_k.cu): the tensor-core Gram with K-major operand tiles must pass
    the very tests the default kernel passes (operator level, K1 Gram, heavy-row half-steps)."""
    env = dict(os.environ)
    env["IALS_WGRAM"] = "kmajor"
    res = subprocess.run([sys.executable, "-m", "pytest", "tests/test_wgram.py", "tests/test_gpu_parity.py",
                          "-m", "gpu", "-q", "-x", "-k",
                          "wgram or gram or half_steps or heavy or c2_full_size"],
                         env=env, cwd=ROOT, timeout=900, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1000:]


def test_kmajor_gram_epochs_match_the_default_kernel(tmp_path):
    """Same products per MMA, same k order; only the partial sums of b are grouped differently."""
    ref = _epochs(tmp_path, "default", {"IALS_WGRAM": ""})
    km = _epochs(tmp_path, "kmajor", {"IALS_WGRAM": "kmajor"})
    for k in ("user", "item"):
        assert np.abs(ref[k] - km[k]).max() <= 2e-5 * np.abs(ref[k]).max()


@pytest.mark.parametrize("threshold", ["2048", "64"])
def test_fused_heavy_rows_pass_the_parity_tests(threshold):
    """IALS_WGRAM=fused (wgram_k.cu, FUSE instantiation): single-job heavy rows are solved in the
    Gram kernel's epilogue.  With the threshold at 64 most rows of the test matrices take that
    route; the oracle comparisons of the CG half-steps and epochs must hold unchanged."""
    env = dict(os.environ)
    env["IALS_WGRAM"] = "fused"
    env["IALS_HEAVY_THRESHOLD"] = threshold
    res = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_golden.py",
                          "-m", "gpu", "-q", "-x", "-k",
                          "half_steps or overfit_cg or c1_config or step_io or empty_rows or golden or c2_full_size"],
                         env=env, cwd=ROOT, timeout=900, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1000:]


def test_fused_heavy_rows_epochs_match_the_default_path(tmp_path):
    ref = _epochs(tmp_path, "default", {"IALS_WGRAM": ""})
    for thr in ("2048", "256"):
        fu = _epochs(tmp_path, f"fused{thr}", {"IALS_WGRAM": "fused", "IALS_HEAVY_THRESHOLD": thr})
        for k in ("user", "item"):  # same systems; the order of the sums inside A p differs
            assert np.abs(ref[k] - fu[k]).max() <= 2e-4 * np.abs(ref[k]).max()


@pytest.mark.parametrize("threads", ["128", "256", "544"])
def test_wide_ialspp_ctas_pass_the_ialspp_parity_tests(threads):
    """IALS_IALSPP_THREADS (cholesky_tile.cu launch_tile): the block solver with wider CTAs -- same
    kernel, only blockDim changes -- must pass the iALS++ parity cases unchanged."""
    env = dict(os.environ)
    env["IALS_IALSPP_THREADS"] = threads
    res = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_golden.py",
                          "-m", "gpu", "-q", "-x", "-k", "ialspp"],
                         env=env, cwd=ROOT, timeout=900, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1000:]


@pytest.mark.parametrize("mode", ["", "tc"])
def test_cholesky_k256_half_epochs_against_the_oracle(mode):
    """K = 256 / 240 Cholesky half-epochs (row stride 256) vs the float64 oracle: the default
    register-tiled kernel and IALS_CHOL=tc (Gram blocks on the tensor cores: wgram_k.cu on both
    row halves + wgram_cross_kernel, cholesky_tile_kernel<2>).  TOL_STEP of test_gpu_parity.py."""
    import json

    env = dict(os.environ)
    env["IALS_CHOL"] = mode
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parity_chol256.py")], env=env, cwd=ROOT,
                         timeout=600, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    for case in ("K256_IALSPP", "K240_ORIGINAL"):
        e = out[case]
        assert e["empty_row_is_zero"]
        for side in (0, 1):
            assert e[f"side{side}_vs_f64"] <= 2e-4 + 2 * e[f"side{side}_oracle32_vs_f64"], (case, e)
