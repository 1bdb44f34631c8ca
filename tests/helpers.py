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


def compare_features(got, ref, evals, eig_order="literal", what="features", ids=None):
    """Asserts |got-ref| <= ATOL + RTOL*|ref| with the conditioning-aware exemptions of SURVEY.md 7.3.

    ``ids`` gives the EFeatureID of every column (default: the 11 columns of compute_features).
    Exempt, and counted in the returned statistics:
      * vector-valued columns (normal, both verticalities) on rows whose relative eigen-gap is
        below 1e-3 -- the eigenvectors are not determined to the tolerance there;
      * the normal's sign on rows with |n_z| < 1e-3 (the z >= 0 canonicalisation is a coin flip);
      * rank-deficient rows (see ``degenerate``): only a loose 1e-2 check in documented order;
      * literal order: rows whose smallest eigenvalue is below float32 covariance resolution get a
        tolerance scaled by the predicted amplification (see the comment in the code).
    """
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.isfinite(got).all(), "%s: non-finite values" % what
    ids = list(range(got.shape[1])) if ids is None else [int(i) for i in ids]
    col = {fid: [c for c, i in enumerate(ids) if i == fid] for fid in set(ids)}
    bad = np.abs(got - ref) > ATOL + RTOL * np.abs(ref)
    ill = ill_conditioned(evals)
    deg = degenerate(evals)
    for fid in (3, 4, 5, 6, 12):
        for c in col.get(fid, []):
            bad[ill, c] = False
    # sign of a normal lying in the xy plane
    if all(f in col for f in (4, 5, 6)):
        cx, cy, cz = col[4][0], col[5][0], col[6][0]
        flat = np.abs(ref[:, cz]) < 1e-3
        for c in (cx, cy, cz):
            alt = np.abs(got[:, c] + ref[:, c]) <= ATOL + RTOL * np.abs(ref[:, c])
            bad[flat, c] &= ~alt[flat]
    if eig_order == "literal":
        # Literal order puts sqrt(lambda_min) in the 1/(s0 + 1e-3) denominator.  A float32 covariance
        # carries ~1e-7 lambda_max of absolute error, i.e. d(s0) ~ 1e-7 lambda_max / (2 s0), which the
        # ratio features amplify by 1/(s0 + 1e-3): rows where that predicted relative error exceeds the
        # tolerance cannot agree between ANY two float32 evaluations (the reference's own included).
        s0 = np.sqrt(np.maximum(evals[:, 0], 1e-300))
        amp = 1e-7 * evals[:, 2] / (2.0 * s0 * (s0 + 1e-3))
        weak = amp > 1e-4
        loose_lit = np.abs(got - ref) > 1e-3 + np.minimum(100.0 * amp, 0.5)[:, None] * (np.abs(ref) + 1.0)
        for fid in (0, 1, 2, 10, 7, 8, 9):
            for c in col.get(fid, []):
                bad[weak, c] = loose_lit[weak, c]
    if eig_order == "documented":
        loose = np.abs(got - ref) > 1e-2 + 1e-2 * np.abs(ref)
        for fid in (3, 4, 5, 6, 12):
            for c in col.get(fid, []):
                loose[:, c] = False
        bad[deg] = loose[deg]
    else:
        bad[deg] = False
    for c in col.get(11, []):          # K_optimal column of compute_features_selected stays 0
        bad[:, c] = got[:, c] != ref[:, c]
    rows = np.unique(np.nonzero(bad)[0])
    assert rows.size == 0, "%s: %d rows out of tolerance, first %s\n got %s\n ref %s\n evals %s" % (
        what, rows.size, rows[:5], got[rows[:3]], ref[rows[:3]], evals[rows[:3]])
    return {"rows": got.shape[0], "ill_conditioned_rows": int(ill.sum()), "degenerate_rows": int(deg.sum())}


def knn_csr(idx):
    n, k = idx.shape
    return np.ascontiguousarray(idx.reshape(-1)).astype(np.uint32), (np.arange(n + 1) * k).astype(np.uint32)


def radius_csr(idx):
    nn_ptr = np.r_[0, (idx >= 0).sum(axis=1).cumsum()].astype(np.uint32)
    return idx[idx >= 0].astype(np.uint32), nn_ptr
