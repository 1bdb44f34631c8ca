"""The optimal-k scan filters candidate sizes with float32 closed-form eigenvalues and widens its too-close-to-call window by
an a-posteriori bound on the entropy error of that form (csrc/features.cu: eigvals3_closed_f32).  This restates the routine in
numpy float32 -- same formulas, same constants -- and checks the bound against float64 eigenvalues on random, planar, linear,
isotropic and nearly coincident spectra.  (CPU only: it tests the mathematics of the bound, the CUDA routine itself is covered
by the k_opt exactness tests in test_gpu_parity.py.)"""
import numpy as np

F = np.float32


def entropy(l0, l1, l2, dtype):
    eps = dtype(1e-3)
    s = l0 + l1 + l2 + eps
    e = [l0 / s, l1 / s, l2 / s]
    return -sum(x * np.log(x + eps) for x in e)                                   # pca.hpp:140-150


def closed_f32(c):
    """c: (n, 6) float32 upper triangle (a00 a01 a02 a11 a12 a22) scaled to max |a_ij| = 1 -> (w (n, 3), err (n,))"""
    a00, a01, a02, a11, a12, a22 = [c[:, i].astype(F) for i in range(6)]
    q = (a00 + a11 + a22) * F(1 / 3)
    b00, b11, b22 = a00 - q, a11 - q, a22 - q
    off2 = a01 * a01 + (a02 * a02 + a12 * a12)
    p2 = b00 * b00 + (b11 * b11 + (b22 * b22 + F(2) * off2))
    ok = p2 > F(1e-30)
    p2s = np.where(ok, p2, F(1))
    ip = (F(1) / np.sqrt(p2s * F(1 / 6))).astype(F)
    p = (p2s * F(1 / 6) * ip).astype(F)
    c00, c11, c22, c01, c02, c12 = b00 * ip, b11 * ip, b22 * ip, a01 * ip, a02 * ip, a12 * ip
    r = F(0.5) * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02))
    r = np.clip(r, F(-1), F(1)).astype(F)
    phi = (np.arccos(r) * F(1 / 3)).astype(F)
    hi = (F(2) * p * np.cos(phi) + q).astype(F)
    lo = (F(2) * p * np.cos(phi + F(2.0943951)) + q).astype(F)
    mid = (F(3) * q - hi - lo).astype(F)
    w = np.maximum(np.stack([lo, mid, hi], 1), F(0))
    dr = F(4e-6)
    dphi = np.minimum(dr / np.sqrt(np.maximum(F(1) - r * r, F(1e-12))), F(2.9e-3)) * F(1 / 3)
    err = F(24) * (F(2) * p * (dphi + F(5e-7)) + F(1e-6))
    qq = np.maximum(q, F(0))
    w = np.where(ok[:, None], w, np.stack([qq, qq, qq], 1))
    err = np.where(ok, err, F(1e-6))
    return w.astype(F), err.astype(F)


def spectra(rng, n):
    """eigenvalue triples covering the hard cases"""
    out = [rng.uniform(0, 1, (n, 3)),                                              # generic
           np.c_[rng.uniform(0.1, 1, n), rng.uniform(0.1, 1, n), np.zeros(n)],      # planes
           np.c_[rng.uniform(0.1, 1, n), np.zeros(n), np.zeros(n)],                 # lines
           np.repeat(rng.uniform(0.1, 1, (n, 1)), 3, 1) * (1 + rng.normal(0, 1e-6, (n, 3))),   # isotropic, nearly coincident
           np.c_[np.ones(n), 10.0 ** rng.uniform(-8, -1, n), 10.0 ** rng.uniform(-8, -1, n)],  # needles: two tiny eigenvalues
           np.c_[np.ones(n), np.ones(n) * (1 + rng.normal(0, 1e-5, n)), 10.0 ** rng.uniform(-6, 0, n)]]  # two large coincident
    return np.abs(np.concatenate(out))


def test_entropy_error_of_the_closed_form_stays_inside_its_bound():
    rng = np.random.default_rng(5)
    lam = spectra(rng, 20000)
    n = len(lam)
    # random rotations
    a = rng.normal(size=(n, 3, 3))
    qm, _ = np.linalg.qr(a)
    cov = np.einsum("nij,nj,nkj->nik", qm, lam, qm)
    cov = 0.5 * (cov + cov.transpose(0, 2, 1))
    scale = np.abs(cov).max(axis=(1, 2))
    cov32 = (cov / scale[:, None, None]).astype(F)                                  # what the kernel feeds the routine
    c6 = np.stack([cov32[:, 0, 0], cov32[:, 0, 1], cov32[:, 0, 2], cov32[:, 1, 1], cov32[:, 1, 2], cov32[:, 2, 2]], 1)
    w32, err = closed_f32(c6)
    w64 = np.maximum(np.linalg.eigvalsh(cov / scale[:, None, None]), 0.0)
    s32 = w32 * scale[:, None].astype(F)
    h32 = entropy(s32[:, 0], s32[:, 1], s32[:, 2], F).astype(np.float64)
    s64 = w64 * scale[:, None]
    h64 = entropy(s64[:, 0], s64[:, 1], s64[:, 2], np.float64)
    diff = np.abs(h32 - h64)
    worst = np.max(diff / err)
    print("closed form: max |h32 - h64| = %.3g, max ratio to the returned bound = %.3f, bound range [%.2g, %.2g]" % (diff.max(), worst, err.min(), err.max()))
    assert (diff <= err).all()
    assert np.median(err[:20000]) < 1e-4                                            # generic spectra: the window stays narrow where the form is well conditioned
