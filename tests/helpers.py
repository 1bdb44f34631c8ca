"""Shared helpers of the parity tests (tolerances of BASELINE.json:north_star)."""
import numpy as np

ATOL, RTOL = 1e-4, 1e-3     # "the 11 features match within 1e-4 absolute / 1e-3 relative"
GAP_MIN = 1e-3              # rows with a smaller relative eigen-gap: vector-valued columns are ill conditioned
VECTOR_COLS = (3, 4, 5, 6)  # VerticalityPGEOF (uses |v0|,|v1|,|v2|) and the normal


def row_eigvals(xyz, nn, nn_ptr, k=None):
    """Ascending covariance eigenvalues (float64) of every CSR row (first k entries if given)."""
    nn_ptr = np.asarray(nn_ptr, np.int64)
    n = len(nn_ptr) - 1
    out = np.zeros((n, 3))
    x64 = np.asarray(xyz, np.float64)
    lens = np.diff(nn_ptr)
    for i in range(n):
        m = lens[i] if k is None else min(k, lens[i])
        if m == 0:
            continue
        p = x64[nn[nn_ptr[i]:nn_ptr[i] + m]]
        c = p - p.mean(0)
        out[i] = np.linalg.eigvalsh(c.T @ c / m)
    return out


def row_eigvals_dense(xyz, idx):
    """Vectorised variant for a dense (n, k) neighbour table."""
    p = np.asarray(xyz, np.float64)[idx]                 # (n, k, 3)
    c = p - p.mean(1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", c, c) / idx.shape[1]
    return np.linalg.eigvalsh(cov)


def ill_conditioned(evals):
    """Rows whose eigenvectors are not determined to 1e-3: relative gap below GAP_MIN."""
    scale = np.maximum(evals[:, 2], 1e-300)
    gap = np.minimum(evals[:, 1] - evals[:, 0], evals[:, 2] - evals[:, 1]) / scale
    return gap < GAP_MIN


def degenerate(evals):
    """Rows that are rank deficient at float32 resolution (lambda_min <= 1e-6 lambda_max).

    sqrt(lambda_min) then carries an absolute error of ~3e-4 sqrt(lambda_max) in ANY float32
    evaluation (the reference's own included): every feature that touches it is outside the
    1e-4 tolerance by construction, and with the literal eigen order -- where the smallest
    eigenvalue sits in the 1/(s0 + 1e-3) denominator -- the error is amplified 1000x.
    """
    return evals[:, 0] <= 1e-6 * np.maximum(evals[:, 2], 1e-300)


def compare_features(got, ref, evals, eig_order="literal", what="features", ids=None, max_weak=None, max_degenerate=None,
                     max_ill=None):
    """Asserts |got-ref| <= ATOL + RTOL*|ref| with the conditioning-aware relaxations of SURVEY.md 7.3.

    ``ids`` gives the EFeatureID of every column (default: the 11 columns of compute_features).
    Every row is checked; what changes with the conditioning of a row is the tolerance, never whether
    it is looked at.  Relaxed, and COUNTED in the returned statistics (callers bound the counts with
    ``max_weak`` / ``max_degenerate`` / ``max_ill``, fractions of the rows, or assert on the returned dict):
      * ``ill_conditioned_rows``: relative eigen-gap below 1e-3 -- the eigenvectors are not determined to
        the tolerance, so the vector-valued columns (normal, both verticalities) only have to be finite
        and inside their range ([-1, 1] / [0, 1]);
      * the normal's sign on rows with |n_z| < 1e-3 (the z >= 0 canonicalisation is a coin flip);
      * ``weak_rows``: rows whose smallest eigenvalue is near float32 covariance resolution get a tolerance
        scaled by the predicted amplification (see the comments in the code);
      * ``degenerate_rows`` (rank deficient at float32 resolution): documented order -> 1e-2 abs/rel;
        literal order -> the amplification-scaled tolerance for the scalar columns and 1e-2 for the
        vector-valued ones (the largest eigenvalue's vector is well determined there).
    """
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.isfinite(got).all(), "%s: non-finite values" % what
    ids = list(range(got.shape[1])) if ids is None else [int(i) for i in ids]
    col = {fid: [c for c, i in enumerate(ids) if i == fid] for fid in set(ids)}
    vec_cols = [c for fid in (3, 4, 5, 6, 12) for c in col.get(fid, [])]
    err = np.abs(got - ref)
    # sign of a normal lying in the xy plane: the z >= 0 canonicalisation is decided by round-off, both signs are right
    if all(f in col for f in (4, 5, 6)):
        flat = np.abs(ref[:, col[6][0]]) < 1e-3
        for c in (col[4][0], col[5][0], col[6][0]):
            err[flat, c] = np.minimum(err[flat, c], np.abs(got[flat, c] + ref[flat, c]))
    bad = err > ATOL + RTOL * np.abs(ref)
    ill = ill_conditioned(evals)
    deg = degenerate(evals)
    weak = np.zeros(len(evals), bool)
    # eigenvectors of a (near-)double eigenvalue: any unit vector of the eigenspace is a correct answer
    for c in vec_cols:
        bad[ill, c] = np.abs(got[ill, c]) > 1.0 + 1e-5
    loose = err > 1e-2 + 1e-2 * np.abs(ref)
    if eig_order == "literal":
        # Literal order puts sqrt(lambda_min) in the 1/(s0 + 1e-3) denominator.  A float32 covariance
        # carries ~1e-7 lambda_max of absolute error, i.e. d(s0) ~ 1e-7 lambda_max / (2 s0), which the
        # ratio features amplify by 1/(s0 + 1e-3): rows where that predicted relative error exceeds the
        # tolerance cannot agree between ANY two float32 evaluations (the reference's own included).
        s0 = np.sqrt(np.maximum(evals[:, 0], 1e-300))
        amp = 1e-7 * evals[:, 2] / (2.0 * s0 * (s0 + 1e-3))
        weak = (amp > 1e-4) & ~deg
        loose_lit = err > 1e-3 + np.minimum(100.0 * amp, 0.5)[:, None] * (np.abs(ref) + 1.0)
        for fid in (0, 1, 2, 10, 7, 8, 9):
            for c in col.get(fid, []):
                bad[weak, c] = loose_lit[weak, c]
        # Rank-deficient rows: the one or two small singular values are pure round-off, anywhere in [0, d] with
        # d = sqrt(4e-7 lambda_max) in a float32 evaluation (0 in float64).  Each column gets the tolerance that follows from
        # its formula: the ratio features divide by (s0 + 1e-3) with s0 <= d; Length = s0; Surface = sqrt(s0 s1 + 1e-6);
        # Volume = cbrt(s0 s1 s2 + 1e-9); Curvature = s2 / (s0 + s1 + s2 + 1e-3).
        lmax = np.maximum(evals[:, 2], 0.0)
        d = np.sqrt(4e-7 * lmax)
        smax_ = np.sqrt(lmax)
        deg_tol = {0: 2e3 * d, 1: 2e3 * d, 2: 2e3 * d, 7: 2.0 * d, 8: 2.0 * d + np.sqrt(d * smax_), 9: np.cbrt(d * lmax),
                   10: 4.0 * d / np.maximum(smax_, 1e-300)}
        for fid, extra_tol in deg_tol.items():
            # the ratio features also scale with 1 / (s0 + 1e-3), s0 anywhere in [0, d]: a relative uncertainty of d / 1e-3
            rel = 1e-2 + (2e3 * d if fid in (0, 1, 2) else 0.0)
            for c in col.get(fid, []):
                bad[deg, c] = (err[:, c] > 1e-3 + rel * np.abs(ref[:, c]) + extra_tol)[deg]
        for c in vec_cols:
            bad[deg & ~ill, c] = loose[deg & ~ill, c]
        # pca.hpp:184 computes VerticalityPGEOF only `if (val0 > 0)`: on a rank-deficient row the literal val0 = sqrt(lambda_min) is
        # round-off, so whether the gate opens is a coin flip in ANY float evaluation -- 0 and the formula's value are both right
        for c in col.get(3, []):
            gate = deg & ((got[:, c] == 0.0) | (ref[:, c] == 0.0)) & (got[:, c] >= 0.0) & (got[:, c] <= 1.0 + 1e-5)
            bad[gate, c] = False
    else:
        # rank-deficient rows, documented order: sqrt(lambda_min) is anywhere in [0, ~5e-4 sqrt(lambda_max)] in a float32
        # evaluation, which Surface = sqrt(s0 s1 + 1e-6) and Volume = cbrt(s0 s1 s2 + 1e-9) turn into up to ~0.03 / ~0.1 s_max
        smax = np.sqrt(np.maximum(evals[:, 2], 0.0))
        extra = {8: 0.03 * smax, 9: 0.1 * smax}
        # nearly rank-deficient rows: d(sqrt(lambda_min)) ~ 1e-7 lambda_max / (2 sqrt(lambda_min)) reaches the tolerance in the
        # columns built on it (planarity, scattering, volume, curvature); scale the tolerance by that prediction
        smin = np.sqrt(np.maximum(evals[:, 0], 1e-300))
        e2 = 1e-7 * evals[:, 2] / (2.0 * smin * (smax + 1e-3))
        weak = (e2 > 1e-5) & ~deg
        for fid in (1, 2, 9, 10):
            for c in col.get(fid, []):
                bad[weak, c] = (err[:, c] > ATOL + RTOL * np.abs(ref[:, c]) + 30.0 * e2 * np.maximum(1.0, smax))[weak]
        for c in range(got.shape[1]):
            if c not in vec_cols:
                bad[deg, c] = (err[:, c] > 1e-2 + 1e-2 * np.abs(ref[:, c]) + extra.get(ids[c], 0.0))[deg]
        for c in vec_cols:
            bad[deg & ~ill, c] = loose[deg & ~ill, c]
    for c in col.get(11, []):          # K_optimal column of compute_features_selected stays 0
        bad[:, c] = got[:, c] != ref[:, c]
    rows = np.unique(np.nonzero(bad)[0])
    assert rows.size == 0, "%s: %d rows out of tolerance, first %s\n got %s\n ref %s\n evals %s" % (
        what, rows.size, rows[:5], got[rows[:3]], ref[rows[:3]], evals[rows[:3]])
    n = max(got.shape[0], 1)
    stats = {"rows": got.shape[0], "ill_conditioned_rows": int(ill.sum()), "degenerate_rows": int(deg.sum()), "weak_rows": int(weak.sum())}
    for name, cap in (("weak_rows", max_weak), ("degenerate_rows", max_degenerate), ("ill_conditioned_rows", max_ill)):
        if cap is not None:
            assert stats[name] <= cap * n, "%s: %d %s of %d exceed the bound %g" % (what, stats[name], name, n, cap)
    return stats


def knn_csr(idx):
    n, k = idx.shape
    return np.ascontiguousarray(idx.reshape(-1)).astype(np.uint32), (np.arange(n + 1) * k).astype(np.uint32)


def radius_csr(idx):
    nn_ptr = np.r_[0, (idx >= 0).sum(axis=1).cumsum()].astype(np.uint32)
    return idx[idx >= 0].astype(np.uint32), nn_ptr
