"""Device self test of the division-free cell lookup used on the hot path: it must be
bit-identical to the IEEE operations the reference's NumPy arithmetic performs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b", [1e-5, 25.850471, 0.02, 0.0371, 8.2e-6, 1.0 / 3.0, 4.88e-4])
def test_fast_cell_lookup_is_bit_identical(b):
    import torch
    from pypic_b200 import _lib, device as D
    dev = D.require_cuda()
    mism = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib.call("pic_dev_selftest_div", float(b), 200_000_000, 7, D.ptr(mism), D.stream())
    assert int(mism.item()) == 0
