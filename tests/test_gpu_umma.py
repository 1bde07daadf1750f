"""tcgen05 plumbing self-test: one CTA, int8 x int8 -> int32 through TMEM, against numpy."""
import numpy as np
import pytest
import torch

from phyloforge_b200._lib import check
from phyloforge_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_in_tmem", [0, 1])
@pytest.mark.parametrize("k", [32, 64, 256])
def test_umma_i8_matches_numpy(engine, k, a_in_tmem):
    rng = np.random.default_rng(k)
    a = rng.integers(-2, 3, size=(128, k), dtype=np.int8)
    b = rng.integers(-2, 3, size=(256, k), dtype=np.int8)
    da, db = torch.from_numpy(a).to(engine.device), torch.from_numpy(b).to(engine.device)
    dc = torch.full((128, 256), -7, dtype=torch.int32, device=engine.device)
    _lib.check_probe(_lib.load_probes().pf_selftest_umma(da.data_ptr(), db.data_ptr(), dc.data_ptr(), k, a_in_tmem,
                                      torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = a.astype(np.int32) @ b.astype(np.int32).T
    assert np.array_equal(dc.cpu().numpy(), want)
