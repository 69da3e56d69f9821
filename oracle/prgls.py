"""CPU oracle for the PR-GLS / coherent-point-drift EM path (NumPy, float64).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package never imports it.

Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py) against outputs of
the unmodified reference functions executed in the build container through oracle/_ref_shim.py and
committed under tests/golden/ (generator: oracle/make_golden.py).

Each function restates the arithmetic of the cited reference lines with broadcasting instead of the
reference's np.tile temporaries; the order of floating-point operations inside each expression is
kept (sum over the 3 coordinates last-axis, row sums along axis 1, ...) so results agree to the last
few ulps.
"""
import numpy as np


# --------------------------------------------------------------------------------------------------
# shared small pieces
# --------------------------------------------------------------------------------------------------
def dist_squares(ref_nx3, tgt_mx3):
    """(M,N) squared distances |ref_n - tgt_m|^2.  Reference: trackerlite.py:361-365."""
    d = ref_nx3[None, :, :] - tgt_mx3[:, None, :]
    return np.sum(np.square(d), axis=2)


def gaussian_kernel(ref_nx3, tgt_mx3, sigma_square):
    """exp(-d^2 / (2 sigma^2)), shape (M,N).  Reference: trackerlite.py:368-372."""
    return np.exp(-dist_squares(ref_nx3, tgt_mx3) / (2 * sigma_square))


def greedy_pairs(corr_mxn, threshold, max_rounds):
    """Greedy one-to-one assignment: repeatedly take the global maximum (first occurrence in
    row-major order, as np.argmax), then zero its row and column.  Stops when max < threshold.
    Reference: track.py:60-70 and trackerlite.py:246-255.  Returns list of (row m, col n)."""
    work = np.array(corr_mxn, copy=True)
    pairs = []
    for _ in range(max_rounds):
        flat = int(np.argmax(work))
        m, n = divmod(flat, work.shape[1])
        if work[m, n] < threshold:
            break
        pairs.append((m, n))
        work[m, :] = 0
        work[:, n] = 0
    return pairs


def prior_track(corr_mxn):
    """Prior used by pr_gls_quick: rows default to 1/N; a greedily matched row becomes
    0.1/(N-1) with 0.9 at the match (threshold 0.5, at most N rounds).  Reference: track.py:58-70."""
    m, n = corr_mxn.shape
    prior = np.ones((m, n)) / n
    for (r, c) in greedy_pairs(corr_mxn, 0.5, n):
        prior[r, :] = 0.1 / (n - 1)
        prior[r, c] = 0.9
    return prior


def simple_match(corr_mxn, threshold=0.1):
    """TrackerLite prior: every entry 0.1/(N-1), greedily matched pairs 0.9; dtype follows the
    input (np.full_like).  Returns (prior, pairs[(ref n, tgt m)]).  Reference: trackerlite.py:242-259."""
    pairs = greedy_pairs(corr_mxn, threshold, corr_mxn.shape[1])
    prior = np.full_like(np.asarray(corr_mxn), 0.1 / (corr_mxn.shape[1] - 1))
    for (r, c) in pairs:
        prior[r, c] = 0.9
    return prior, np.array([(c, r) for (r, c) in pairs])


# --------------------------------------------------------------------------------------------------
# Tracker flavour: pr_gls_quick
# --------------------------------------------------------------------------------------------------
def pr_gls_quick(X, Y, corr, BETA=300, max_iteration=20, LAMBDA=0.1, vol=1E8):
    """PR-GLS EM of the U-Net workflow.  Reference: track.py:11-114.

    E-step  (track.py:81-88):  P = prior*exp(-d^2/2s2) / (rowsum + g*(2 pi s2)^1.5/((1-g)*vol))
    M-step  (track.py:91-97):  (G diag(p) + L s2 I)^T C^T = (Y^T P - X^T diag(p))^T
    update  (track.py:100-112): T_X = X + (C G)^T ; g = 1 - sum(P)/M ; s2 = sum(P d^2)/(3 sum P), >= 1
    """
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    n, m = X.shape[0], Y.shape[0]
    gamma = 0.1
    gram = np.exp(-dist_squares(X, X) / (2 * BETA * BETA))
    C = np.zeros((3, n))
    sigma_square = np.sum(dist_squares(X, Y)) / (3 * n * m)
    prior = prior_track(corr)
    T_X = X.copy()
    P = None
    for _ in range(1, max_iteration):
        d2 = dist_squares(T_X, Y)
        P1 = prior * np.exp(-d2 / (2 * sigma_square))
        den = np.sum(P1, axis=1) + gamma * (2 * np.pi * sigma_square) ** 1.5 / ((1 - gamma) * vol)
        P = P1 / den[:, None]
        p = np.sum(P, axis=0)
        a = gram * p[None, :] + LAMBDA * sigma_square * np.identity(n)
        b = Y.T @ P - X.T * p[None, :]
        C = np.linalg.solve(a.T, b.T).T
        T_X = X + (C @ gram).T
        M_P = np.sum(P)
        gamma = 1 - M_P / m
        d2 = dist_squares(T_X, Y)
        sigma_square = np.sum(P * d2) / (3 * M_P)
        if sigma_square < 1:
            sigma_square = 1
    return P, T_X, C


def predict_one_rep(pred_pre_lx3, inter_nx3, beta, C_3xn):
    """Apply one fitted transform to the tracked cells.  Reference: tracker.py:1269-1289.
    G[n,l] = exp(-|pred_l - inter_n|^2 / 2 beta^2);  post = pre + (C G)^T."""
    g = np.exp(-dist_squares(pred_pre_lx3, inter_nx3) / (2 * beta * beta))  # (N, L)
    return pred_pre_lx3 + (C_3xn @ g).T


# --------------------------------------------------------------------------------------------------
# TrackerLite flavour
# --------------------------------------------------------------------------------------------------
def estimate_posterior(prior_mxn, sigma_square, pred_ref_nx3, tgt_mx3, ratio_outliers, vol=1):
    """Reference: trackerlite.py:375-382."""
    like = gaussian_kernel(pred_ref_nx3, tgt_mx3, sigma_square)
    joint = (1 - ratio_outliers) * prior_mxn * like / (2 * np.pi * sigma_square) ** 1.5
    den = np.sum(joint, axis=1) + ratio_outliers / vol
    return joint / den[:, None]


def solve_movements_ref(sigma_square, lambda_, post_mxn, ref_nx3, tgt_mx3, gram_nxn):
    """Reference: trackerlite.py:409-417."""
    n = ref_nx3.shape[0]
    p = np.sum(post_mxn, axis=0)
    coef = gram_nxn * p[None, :] + lambda_ * sigma_square * np.identity(n)
    dep = tgt_mx3.T @ post_mxn - ref_nx3.T * p[None, :]
    return np.linalg.solve(coef.T, dep.T).T


def prgls_with_two_ref(init_match_mxn, ptrs_tgt_mx3, prts_ref_nx3, tracked_ref_lx3, beta, lambda_,
                       max_iteration=2000, return_iterations=False):
    """Reference: trackerlite.py:309-358.  M-step is relative to the *current* prediction; the first
    increment is discarded; gamma >= 1e-4; stop when the Frobenius norm of the increment < 1e-3."""
    gram_nn = gaussian_kernel(prts_ref_nx3, prts_ref_nx3, beta ** 2)
    gram_nl = gaussian_kernel(tracked_ref_lx3, prts_ref_nx3, beta ** 2)   # (N, L)
    ratio_outliers = 0.05
    sigma_square = dist_squares(prts_ref_nx3, ptrs_tgt_mx3).mean() / 3
    pred_n = np.array(prts_ref_nx3, dtype=np.float64, copy=True)
    pred_l = np.array(tracked_ref_lx3, dtype=np.float64, copy=True)
    post = None
    its = 0
    for iteration in range(1, max_iteration):
        its = iteration
        post = estimate_posterior(init_match_mxn, sigma_square, pred_n, ptrs_tgt_mx3, ratio_outliers)
        C = solve_movements_ref(sigma_square, lambda_, post, pred_n, ptrs_tgt_mx3, gram_nn)
        move_n = (C @ gram_nn).T
        move_l = (C @ gram_nl).T
        if iteration > 1:
            pred_n += move_n
            pred_l += move_l
        s = np.sum(post)
        ratio_outliers = 1 - s / ptrs_tgt_mx3.shape[0]
        if ratio_outliers < 1E-4:
            ratio_outliers = 1E-4
        sigma_square = np.sum(dist_squares(pred_n, ptrs_tgt_mx3) * post) / (3 * s)
        if np.sqrt(np.sum(np.square(move_n))) < 1E-3:
            break
    if return_iterations:
        return pred_l, post, its
    return pred_l, post


def prgls_quick(init_match_mxn, ptrs_tgt_mx3, tracked_ref_nx3, beta, lambda_, max_iteration=2000):
    """Reference: trackerlite.py:262-306 (the L == N special case of prgls_with_two_ref)."""
    pred, post = prgls_with_two_ref(init_match_mxn, ptrs_tgt_mx3, tracked_ref_nx3, tracked_ref_nx3,
                                    beta, lambda_, max_iteration)
    return pred, post


# --------------------------------------------------------------------------------------------------
# ensemble scheduling + reduce
# --------------------------------------------------------------------------------------------------
def get_remote_vols(ensemble, vol):
    """Reference: track.py:605-610."""
    interval = (vol - 1) // ensemble
    start = (vol - 1) % ensemble + 1
    return list(range(start, vol - interval + 1, interval))


def get_reference_vols(ensemble, vol, adjacent=False):
    """Reference: track.py:575-602."""
    if not ensemble:
        return [vol - 1]
    if vol - 1 < ensemble:
        return list(range(1, vol))
    if adjacent:
        return list(range(vol - ensemble, vol))
    return get_remote_vols(ensemble, vol)


def get_volumes_list(current_vol, skip_volumes, sampling_number=20, adjacent=False, start_vol=1):
    """Reference: trackerlite.py:420-438."""
    assert current_vol > start_vol
    if current_vol - start_vol < sampling_number:
        vols = list(range(start_vol, current_vol))
    elif adjacent:
        vols = list(range(current_vol - sampling_number, current_vol))
    else:
        interval = (current_vol - start_vol) // sampling_number
        start = (current_vol - start_vol) % sampling_number + start_vol
        vols = list(range(start, current_vol - interval + 1, interval))
    return [v for v in vols if v not in skip_volumes]


def trim_mean(stack_exl3, proportiontocut=0.1):
    """scipy.stats.trim_mean(..., axis=0) restated: sort along axis 0, drop int(p*E) from each end,
    average the rest.  Call sites: tracker.py:1507, trackerlite.py:123."""
    a = np.sort(np.asarray(stack_exl3, dtype=np.float64), axis=0)
    e = a.shape[0]
    cut = int(proportiontocut * e)
    return np.mean(a[cut:e - cut], axis=0)
