"""vlb_diag_cache_peaks: the measured cache ceilings bench.py holds the traversal kernel against are sane on a B200."""
import pytest

pytestmark = pytest.mark.gpu


def test_cache_peaks_are_ordered_and_plausible(ctx):
    pk = ctx.cache_peaks()
    # an L1-resident stream beats an L2-resident one, which beats HBM (~6.5 TB/s measured copy peak); scattered 16-byte loads
    # (8 lines per request) deliver less than the coalesced L1 stream
    assert pk["l1_read_gbs"] > pk["l2_read_gbs"] > 5000.0
    assert pk["l1_scatter_lines_per_request"] == 8
    assert 0 < pk["l1_scatter_gbs"] < pk["l1_read_gbs"] * 1.05
    assert abs(pk["l1_scatter_wavefronts_per_s"] - 8 * pk["l1_scatter_requests_per_s"]) <= 1e-6 * pk["l1_scatter_wavefronts_per_s"]
