"""-m gpu: the reference's test-time path (model.py:107-446) on the CUDA kernels against the reference's golden run
(tests/golden/renet_eval_tiny.npz, written by oracle/gen_golden.py from the unmodified reference)."""
import pytest

from helpers import check_eval_against_golden, eval_flow, eval_setup

pytestmark = pytest.mark.gpu


def test_eval_flow_matches_reference_golden():
    from renet_b200 import _lib
    ctx = eval_setup('cuda:0')
    n0 = _lib.launch_count()
    res = eval_flow(ctx, 'cuda:0')
    assert _lib.launch_count() > n0                        # the kernels ran, not a fallback
    check_eval_against_golden(res, ctx['ev'])
