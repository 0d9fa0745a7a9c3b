"""-m gpu: every kernel family through the C ABI against the oracle (tests/gpu_checks.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(fn):
    rows = fn()
    bad = [(n, e, t) for (n, e, t) in rows if not (e <= t)]
    assert not bad, "\n".join(f"{n}: err {e:.3e} > tol {t:.1e}" for n, e, t in bad)
    assert rows


@pytest.mark.parametrize("name", ["check_codebook_and_mask", "check_gate", "check_pool", "check_conv_fwd",
                                  "check_conv_epilogues", "check_conv_wgrad", "check_thin", "check_conv_variants", "check_sn", "check_conv_ex", "check_conv_tf32",
                                  "check_optim", "check_flow", "check_conv_bnstats",
                                  "check_batched_paths", "check_conv_cluster"])
def test_kernel_family(name):
    assert torch.cuda.is_available()
    import gpu_checks
    _run(getattr(gpu_checks, name))
