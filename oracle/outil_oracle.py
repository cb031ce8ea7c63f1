"""Oracle (test infrastructure) for the reference's ``utils/outil.py`` hot-path functions.

numpy restatement with a *defined* fp32 operation order so that the CUDA
kernels can be compared bit-for-bit.  Paths cite ``/root/reference``.
"""
import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------
# utils/outil.py:21-29  getWHTensor / getWHTensor_Int
# --------------------------------------------------------------------------
def getWHTensor(h, w):
    """Cell-centre coordinates of an (h, w) feature grid, flattened row-major.

    utils/outil.py:21-24.  "W" runs over dim 2 (rows / y), "H" over dim 3
    (columns / x) - the reference's swapped naming is kept.
    """
    r = np.arange(h, dtype=f32).reshape(-1, 1).repeat(w, 1).reshape(-1)
    c = np.arange(w, dtype=f32).reshape(1, -1).repeat(h, 0).reshape(-1)
    W = (r + f32(0.5)) / f32(h)
    H = (c + f32(0.5)) / f32(w)
    return (W - f32(0.5)) * f32(2), (H - f32(0.5)) * f32(2)


def getWHTensor_Int(h, w):
    """utils/outil.py:26-29."""
    r = np.arange(h, dtype=np.int64).reshape(-1, 1).repeat(w, 1).reshape(-1)
    c = np.arange(w, dtype=np.int64).reshape(1, -1).repeat(h, 0).reshape(-1)
    return r, c


# --------------------------------------------------------------------------
# utils/outil.py:32-45  mutualMatching
# --------------------------------------------------------------------------
def mutualMatching(featA, featB, return_score=False):
    """Mutual nearest neighbours with non-zero score.

    featA (C, NA), featB (C, NB) fp32.  A pair (i, j) survives iff
    i = argmax_i' S[i', j], j = argmax_j' S[i, j'] and S[i, j] * S[i, j] > 0
    (fp32 product, utils/outil.py:36-43).  Output sorted by i (``nonzero()``
    row-major order).  Ties: first index (the reference's ``topk`` tie-break is
    implementation-defined).
    """
    featA = np.ascontiguousarray(featA, dtype=f32)
    featB = np.ascontiguousarray(featB, dtype=f32)
    score = featA.T @ featB                       # (NA, NB) fp32 sgemm
    col_arg = score.argmax(axis=0)                # (NB,)  best source per target
    row_arg = score.argmax(axis=1)                # (NA,)  best target per source
    i = np.arange(score.shape[0])
    j = row_arg
    v = score[i, j]
    keep = (col_arg[j] == i) & ((v * v) > 0)
    idx1 = i[keep].astype(np.int64)
    idx2 = j[keep].astype(np.int64)
    if return_score:
        return idx1, idx2, score
    return idx1, idx2


# --------------------------------------------------------------------------
# utils/outil.py:68-87  Homography (4-point DLT through LAPACK SVD)
# --------------------------------------------------------------------------
def dlt_matrix(X, Y):
    """(N,4,3) fp32 X (source), Y (target) -> A (N,8,9) fp64.

    Entries are fp32 products upcast to fp64 (utils/outil.py:73-81: the
    products ``v_ * u`` are numpy float32 before they are stored in the
    float64 array)."""
    X = np.asarray(X, dtype=f32)
    Y = np.asarray(Y, dtype=f32)
    N = X.shape[0]
    A = np.zeros((N, 8, 9))
    one = np.ones(N)
    zero = np.zeros(N)
    for i in range(4):
        u, v, u_, v_ = Y[:, i, 0], Y[:, i, 1], X[:, i, 0], X[:, i, 1]
        A[:, 2 * i] = np.stack([zero, zero, zero, -u, -v, -one, v_ * u, v_ * v, v_], axis=1)
        A[:, 2 * i + 1] = np.stack([u, v, one, zero, zero, zero, -u_ * u, -u_ * v, -u_], axis=1)
    return A


def Homography(X, Y):
    """utils/outil.py:68-87.  Returns (N,3,3) fp32, unit Frobenius norm, LAPACK's sign."""
    A = dlt_matrix(X, Y)
    _, _, vh = np.linalg.svd(A)
    return vh[:, 8].reshape(-1, 3, 3).astype(f32)


def householder_null_vector(A):
    """Null vector of one 8x9 fp64 matrix as LAPACK's dgesdd returns it in
    ``Vh[8]``: the unblocked lower-bidiagonalisation (dgebd2, m < n) builds
    right reflectors G_1..G_8 and ``Vh[8] = (G_1 G_2 ... G_8 e_9)^T``
    (SURVEY.md section 7 step 4 / A.3 #6).  This is the algorithm the CUDA
    kernel implements; kept here (pure python, small cases only) so the kernel
    can be cross-checked without LAPACK in the loop."""
    A = np.array(A, dtype=np.float64)
    m, n = A.shape
    taus, vs = [], []
    for i in range(m):
        # right reflector G_i annihilates A[i, i+1:]
        alpha = A[i, i]
        x = A[i, i + 1:].copy()
        xnorm = np.sqrt(np.sum(x * x))
        if xnorm == 0.0:
            tau = 0.0
            v = np.zeros(n - i)
            v[0] = 1.0
        else:
            beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
            tau = (beta - alpha) / beta
            v = np.concatenate([[1.0], x / (alpha - beta)])
            A[i, i] = beta
            A[i, i + 1:] = 0.0
        taus.append(tau)
        vs.append(v)
        if tau != 0.0 and i + 1 < m:
            w = A[i + 1:, i:] @ v
            A[i + 1:, i:] -= tau * np.outer(w, v)
        # left reflector H_i annihilates A[i+2:, i]
        if i + 1 < m:
            alpha = A[i + 1, i]
            x = A[i + 2:, i].copy()
            xnorm = np.sqrt(np.sum(x * x))
            if xnorm != 0.0:
                beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
                tauq = (beta - alpha) / beta
                u = np.concatenate([[1.0], x / (alpha - beta)])
                A[i + 1, i] = beta
                A[i + 2:, i] = 0.0
                w = u @ A[i + 1:, i + 1:]
                A[i + 1:, i + 1:] -= tauq * np.outer(u, w)
    h = np.zeros(n)
    h[n - 1] = 1.0
    for i in range(m - 1, -1, -1):
        v = vs[i]
        h[i:] -= taus[i] * v * (v @ h[i:])
    return h


# --------------------------------------------------------------------------
# utils/outil.py:97-113  Prediction / ScoreRANSAC
# --------------------------------------------------------------------------
def Prediction(X, Y, H21):
    """Reprojection error, utils/outil.py:97-100.

    X, Y (M,3) fp32; H21 (N,3,3) fp32 -> (N,M) fp32.  Defined fp32 order (no
    FMA): est_k = (Y0*Hk0 + Y1*Hk1) + Y2*Hk2;  ex = est_0/est_2, ey = est_1/est_2;
    err = sqrt((X0-ex)^2 + (X1-ey)^2)."""
    X = np.asarray(X, dtype=f32)
    Y = np.asarray(Y, dtype=f32)
    H = np.asarray(H21, dtype=f32)
    y0, y1, y2 = Y[None, :, 0], Y[None, :, 1], Y[None, :, 2]
    with np.errstate(all="ignore"):
        e = [(y0 * H[:, k, 0, None] + y1 * H[:, k, 1, None]) + y2 * H[:, k, 2, None] for k in range(3)]
        ex = e[0] / e[2]
        ey = e[1] / e[2]
        dx = X[None, :, 0] - ex
        dy = X[None, :, 1] - ey
        return np.sqrt(dx * dx + dy * dy)


def det3(H):
    """fp32 determinant of (N,3,3) by partial-pivoting LU, no FMA (stand-in for
    ``torch.det``, utils/outil.py:108; only the comparison with 1e-6 is used)."""
    H = np.array(H, dtype=f32)
    N = H.shape[0]
    out = np.zeros(N, dtype=f32)
    for n in range(N):
        a = H[n].copy()
        sign = f32(1)
        p = int(np.argmax(np.abs(a[:, 0])))
        if a[p, 0] == 0:
            out[n] = 0
            continue
        if p != 0:
            a[[0, p]] = a[[p, 0]]
            sign = -sign
        l1 = a[1, 0] / a[0, 0]
        l2 = a[2, 0] / a[0, 0]
        a11 = a[1, 1] - l1 * a[0, 1]
        a12 = a[1, 2] - l1 * a[0, 2]
        a21 = a[2, 1] - l2 * a[0, 1]
        a22 = a[2, 2] - l2 * a[0, 2]
        if abs(a21) > abs(a11):
            a11, a12, a21, a22 = a21, a22, a11, a12
            sign = -sign
        if a11 == 0:
            out[n] = 0
            continue
        l = a21 / a11
        u22 = a22 - l * a12
        out[n] = sign * ((a[0, 0] * a11) * u22)
    return out


def ScoreRANSAC(match1, match2, tolerance, samples, det_fn=det3):
    """utils/outil.py:102-113 -> (H21 (N,3,3) fp32, gated inlier counts (N,) int64)."""
    X = match1[samples]
    Y = match2[samples]
    H21 = Homography(X, Y)
    dets = det_fn(H21)
    err = Prediction(match1, match2, H21)
    inl = err < f32(tolerance)
    return H21, inl.sum(axis=1).astype(np.int64) * (dets > f32(1e-6)).astype(np.int64)


# --------------------------------------------------------------------------
# utils/outil.py:117-164  RANSAC
# --------------------------------------------------------------------------
def unique_samples(samples):
    """utils/outil.py:123-133: drop (not redraw) rows with any repeated index."""
    s = np.asarray(samples)
    dup = ((s[:, 0] == s[:, 1]) | (s[:, 0] == s[:, 2]) | (s[:, 0] == s[:, 3]) |
           (s[:, 1] == s[:, 2]) | (s[:, 1] == s[:, 3]) | (s[:, 2] == s[:, 3]))
    return s[~dup]


def RANSAC_from_samples(match1, match2, samples, tolerance, nbMaxIter=100, det_fn=det3):
    """``outil.RANSAC`` (utils/outil.py:117-164) with the ``torch.randint`` draw
    (line 120) hoisted out: ``samples`` is the (nbIter,4) integer array that call
    returned.  Same chunk-of-100 semantics, the zero-inlier-chunk early return
    (:145-146) and the remainder chunk without that check (:153-160).

    Returns (H (3,3) fp32, nbInlier int64, isInlier (M,) bool, match2[isInlier])
    or (None, 0, [], []).  Raises TypeError where the reference does
    (``bestParams[None]`` with bestParams None, :162)."""
    match1 = np.asarray(match1, dtype=f32)
    match2 = np.asarray(match2, dtype=f32)
    us = unique_samples(samples)
    nbLoop = len(us) // nbMaxIter
    bestParams, bestInlier = None, 0
    for i in range(nbLoop):
        H21, nbInlier = ScoreRANSAC(match1, match2, tolerance, us[i * nbMaxIter:(i + 1) * nbMaxIter], det_fn)
        best = int(np.argmax(nbInlier))
        if nbInlier[best] == 0:
            return None, 0, [], []
        elif nbInlier[best] > bestInlier:
            bestParams = H21[best]
            bestInlier = nbInlier[best]
    if len(us) - nbLoop * nbMaxIter > 0:
        H21, nbInlier = ScoreRANSAC(match1, match2, tolerance, us[nbLoop * nbMaxIter:], det_fn)
        best = int(np.argmax(nbInlier))
        if nbInlier[best] > bestInlier:
            bestParams = H21[best]
            bestInlier = nbInlier[best]
    if bestParams is None:
        raise TypeError("'NoneType' object is not subscriptable")
    err = Prediction(match1, match2, bestParams[None])[0]
    isInlier = err < f32(tolerance)
    return bestParams, np.int64(bestInlier), isInlier, match2[isInlier]


def RANSAC(nbIter, match1, match2, tolerance, nbPoint=4, seed=None):
    """Convenience wrapper drawing samples on the CPU generator (the reference
    draws them on the CUDA generator, so streams differ; parity tests pass the
    same ``samples`` to both sides instead)."""
    import torch
    if seed is not None:
        torch.manual_seed(seed)
    samples = torch.randint(len(match1), (nbIter, nbPoint)).numpy()
    return RANSAC_from_samples(match1, match2, samples, tolerance)


def inlier_mask_grid(Wt, Ht, index2, isInlier, h16, w16):
    """quick_start/coarseAlignFeatMatch.py:166-173 InlierMask on the target grid."""
    idx = np.asarray(index2)[np.asarray(isInlier, dtype=bool)]
    m = np.zeros((h16, w16), dtype=f32)
    r = ((Wt[idx] / f32(2) + f32(0.5)) * f32(h16)).astype(np.int64)
    c = ((Ht[idx] / f32(2) + f32(0.5)) * f32(w16)).astype(np.int64)
    m[r, c] = 1
    return m
