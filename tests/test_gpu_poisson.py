"""The Poisson draw of the diagonal update (csrc/lq_k1.cuh k1_poisson: one inverse-CDF draw per bond
and window) against the reference's sampler (looper/poisson_distribution.h:44-113) through its own
golden, test/poisson_distribution.op: MEAN = 3, COUNT = 2^20, rows "r poisson frequency error".
The generators differ (Philox vs mt19937), so the comparison is statistical: every bin of the GPU
histogram within 4.5 sigma of the exact Poisson weight AND of the golden's own frequency (combined
error), plus a chi-square over the bins."""
import ctypes as C

import numpy as np
import pytest

from test_oracle import POISSON_OP

pytestmark = pytest.mark.gpu


def test_k1_poisson_draw_matches_reference_golden():
    import looper_b200 as lq
    mean, count, nbins = 3.0, 1 << 20, 15
    lq.lib.lq_debug_poisson.argtypes = [C.c_double, C.c_longlong, C.c_ulonglong, C.c_void_p, C.c_int]
    hist = np.zeros(nbins, dtype=np.uint64)
    assert lq.lib.lq_debug_poisson(mean, count, 29833, hist.ctypes.data, nbins) == 0
    hist = hist.astype(np.float64)
    rows = np.array([[float(x) for x in ln.split()] for ln in POISSON_OP.strip().splitlines()])
    assert rows.shape == (nbins, 4)
    exact = np.exp(-mean) * np.cumprod(np.concatenate([[1.0], mean / np.arange(1, nbins)]))
    assert np.allclose(rows[:, 1], exact, rtol=6e-3)            # the golden's own exact column (3 digits)
    freq = hist / count
    err = np.sqrt(np.maximum(exact * count, 1.0)) / count
    assert hist.sum() > count * (1 - 1e-5)                      # P(K >= 15 | mean 3) = 6e-7
    for r in range(nbins):
        assert abs(freq[r] - exact[r]) < 4.5 * err[r], (r, freq[r], exact[r], err[r])
        # golden frequency: 3 printed digits -> half a unit of the last digit joins its error
        gerr = np.hypot(np.hypot(err[r], rows[r, 3]), 0.5 * 10 ** (np.floor(np.log10(rows[r, 2])) - 2))
        assert abs(freq[r] - rows[r, 2]) < 4.5 * gerr, (r, freq[r], rows[r, 2], gerr)
    sel = exact * count > 20
    chi2 = (((hist - exact * count) ** 2) / (exact * count))[sel].sum()
    assert chi2 < 45.0, chi2                                     # 13 bins: P(chi2 > 45) ~ 2e-5
