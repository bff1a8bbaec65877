"""Shared comparators implementing the parity definitions of SURVEY.md section 8c / BASELINE.json north_star."""
import numpy as np

REL = 1e-5  # float tolerance stated by north_star: distances within 1e-5 relative


def assert_topk_exact(dist, ids, gdist, gids, count=None, gcount=None):
    """Integer-valued data: (dist,id) lists bit-identical."""
    assert np.array_equal(ids, gids), f"{(ids != gids).sum()} ids differ"
    assert np.array_equal(np.asarray(dist, np.float32).view(np.uint32), np.asarray(gdist, np.float32).view(np.uint32))
    if gcount is not None:
        assert np.array_equal(count, gcount)


def assert_topk_close(dist, ids, gdist, gids, rtol=REL, atol=1e-6):
    """Float data: distances within rtol; id lists equal except where the differing entries are ties inside the
    tolerance (a swap of two near-equal neighbours, or a near-tie with the kth distance at the list's end)."""
    dist, gdist = np.asarray(dist, np.float64), np.asarray(gdist, np.float64)
    fin = np.isfinite(gdist)
    assert np.array_equal(fin, np.isfinite(dist))
    with np.errstate(invalid="ignore"):
        tol = rtol * np.abs(gdist) + atol
        err = np.abs(dist - gdist)
    assert (err[fin] <= tol[fin]).all(), "distance outside tolerance"
    for q, r in np.argwhere((ids != gids) & fin):
        want = gids[q, r]
        where = np.nonzero(ids[q] == want)[0]
        if len(where):  # present at another rank: must be a swap inside the tolerance
            r2 = where[0]
            assert abs(gdist[q, r2] - gdist[q, r]) <= 2 * tol[q, r], f"query {q}: order differs outside tolerance"
        else:  # dropped: only allowed for a near-tie with the kth distance
            last = np.nonzero(fin[q])[0][-1]
            assert abs(gdist[q, r] - gdist[q, last]) <= 2 * tol[q, r], f"query {q}: id set differs outside tolerance"
