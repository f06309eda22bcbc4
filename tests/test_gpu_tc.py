"""GPU: the tcgen05 / tensor-memory conventions (csrc/bgmm_tc.cuh) — D = A . B^T with kind::tf32 for every operand layout."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tf32(a):
    return (np.ascontiguousarray(a, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


# K-major operands only: with either operand MN-major in the no-swizzle layout the B200 returned an all-zero tile for
# kind::tf32 (measured in round 2), so the fp32-mode kernels stage every operand K-major.
# a_mn = 2: the A operand in tensor memory (tcgen05.st by the thread that owns the row, tcgen05.mma [d], [a], b-desc)
# b_mn = 3: B K-major in the 128-byte-swizzle layout (what the statistics kernel stages), needs Kd % 32 == 0
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (2, 0), (2, 3)])
@pytest.mark.parametrize("n,kd", [(16, 8), (64, 24), (256, 32), (32, 64), (160, 64)])
def test_tcgen05_tf32_matmul(n, kd, a_mn, b_mn):
    import torch
    if b_mn == 3 and kd % 32:
        pytest.skip("swizzled rows hold 32 k")
    from bayesml_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n + kd)
    a = rng.normal(size=(128, kd)).astype(np.float32)
    b = rng.normal(size=(n, kd)).astype(np.float32)
    ad, bd = torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda()
    dd = torch.zeros((128, n), dtype=torch.float32, device="cuda")
    _lib.check(lib.bgmm_tc_selftest(ad.data_ptr(), bd.data_ptr(), dd.data_ptr(), n, kd, a_mn, b_mn,
                                    torch.cuda.current_stream().cuda_stream), "bgmm_tc_selftest")
    torch.cuda.synchronize()
    want = _tf32(a).astype(np.float64) @ _tf32(b).astype(np.float64).T       # the tensor core reads the top 19 bits
    got = dd.cpu().numpy()
    assert np.allclose(got, want, rtol=0, atol=2e-5 * np.sqrt(kd)), np.max(np.abs(got - want))


def test_tcgen05_round_trip_latency():
    """Issue -> execute -> tcgen05.commit -> mbarrier arrival observed by the issuing thread, one 128 x N x 8 tf32 MMA at a time:
    the latency that bounds a pipeline with few stages in flight.  Reported, and sanity-bounded."""
    import torch
    from bayesml_b200 import _lib
    lib = _lib.load()
    for mode, what in ((4, "round trip, one MMA per commit"), (5, "sixteen MMAs per commit, cycles per group")):
        out = {}
        for n in (16, 32, 64, 96, 128, 144, 160, 192, 224, 256):
            a = torch.zeros((128, 8), dtype=torch.float32, device="cuda")
            b = torch.zeros((n, 8), dtype=torch.float32, device="cuda")
            d = torch.zeros((128, n), dtype=torch.float32, device="cuda")
            _lib.check(lib.bgmm_tc_selftest(a.data_ptr(), b.data_ptr(), d.data_ptr(), n, 8, 2, mode,
                                            torch.cuda.current_stream().cuda_stream), "bgmm_tc_selftest")
            torch.cuda.synchronize()
            out[n] = float(d[0, 0].item())
        print(f"tcgen05 kind::tf32 M=128 K=8, A in tensor memory ({what}):", out)
        assert all(50.0 < v < 50000.0 for v in out.values()), out
