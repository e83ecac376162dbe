"""examples/cg_order.cpp on a GPU: the C++ host (XTC file -> gorder_gpu_run_xtc -> gorder_results_*) prints the table the Python
mirror computes from the same frames.  The non-GPU half of this check (it builds, and refuses to run without a device) is
tests/test_example_cpu.py."""
import pytest

from test_example_cpu import test_cpp_host_builds_and_needs_a_gpu as _run_example

pytestmark = pytest.mark.gpu


def test_cpp_host_on_the_gpu(tmp_path):
    import torch
    assert torch.cuda.is_available()
    _run_example(tmp_path)


def test_tpr_host_on_the_gpu(tmp_path):
    """run file + trajectory -> the reference's cg_order_asymmetric_errors.yaml, through C++ only (examples/tpr_order.cpp)."""
    from test_example_cpu import test_tpr_host_builds_and_needs_a_gpu as run
    import torch
    assert torch.cuda.is_available()
    run(tmp_path)
