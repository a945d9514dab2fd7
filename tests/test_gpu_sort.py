"""GPU parity: own radix sort / scans vs torch (integer work: bit-exact)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 31, 4096, 4097, 100_003, 3_000_017])
@pytest.mark.parametrize("bits", [(0, 45), (0, 8), (3, 21)])
def test_radix_sort_pairs(n, bits):
    from emd_b200 import raster_ops as R
    g = torch.Generator().manual_seed(n + bits[1])
    keys = torch.randint(0, 2 ** 46, (n,), generator=g, dtype=torch.int64)
    if n > 10:
        keys[: n // 3] = keys[n // 3: 2 * (n // 3)][: n // 3]  # plenty of duplicates -> stability matters
    vals = torch.arange(n, dtype=torch.int32)
    lo, hi = bits
    sub = (keys >> lo) & ((1 << (hi - lo)) - 1)
    order = torch.sort(sub, stable=True).indices
    k, v = R.radix_sort_pairs(keys.cuda(), vals.cuda(), lo, hi)
    assert torch.equal(k.cpu(), keys[order])
    assert torch.equal(v.cpu(), vals[order])


@pytest.mark.parametrize("n", [1, 5, 4096, 4097, 1_000_003])
def test_cumsum(n):
    from emd_b200 import raster_ops as R
    g = torch.Generator().manual_seed(n)
    x = torch.randint(0, 50, (n,), generator=g, dtype=torch.int32)
    cum, total = R.cumsum_tiles(x.cuda())
    ref = torch.cumsum(x.to(torch.int64), 0)
    assert torch.equal(cum.cpu(), ref) and total == int(ref[-1])
