"""Hoeffding-Bentkus upper confidence bound - the host-side scalar step of the RCPS sweep.

Same names, arguments and return values as the reference's ``core/calibration/bounds.py`` (h1 :6, hoeffding_plus :10,
bentkus_plus :13, HB_mu_plus :17-29).  It stays float64 Python + scipy on the host, exactly like the reference: it is
one scalar root-find per lambda step and never touches a pixel.  ``WSR_mu_plus`` (:31-42) has no caller in the
reference and is out of scope.

On top of the reference API, :func:`hb_stop_bracket` caches, per (n, alpha, delta), the level set of the bound
``{muhat : HB_mu_plus(muhat) > alpha}`` so that the sweep does not need one brentq solve per lambda column.
"""
import functools
import math
import warnings

import numpy as np
from scipy.optimize import brentq
from scipy.stats import binom


def h1(y, mu):
    """Binary relative entropy KL(y || mu)."""
    return y * np.log(y / mu) + (1 - y) * np.log((1 - y) / (1 - mu))


def hoeffding_plus(mu, x, n):
    """Log of Hoeffding's tail bound for an empirical mean x of n draws with true mean mu."""
    return -n * h1(np.minimum(mu, x), mu)


def bentkus_plus(mu, x, n):
    """Log of Bentkus' tail bound (e * binomial cdf, floored at 1e-10 before the log)."""
    return np.log(max(binom.cdf(np.floor(n * x), n, mu), 1e-10)) + 1


def HB_mu_plus(muhat, n, delta, maxiters=1000):
    """Upper confidence bound on a [0,1] mean from ``muhat`` over ``n`` samples at level ``delta``.

    Root in mu of min(hoeffding, bentkus) - log(delta) on [muhat, 1-1e-10]; 1 when even mu -> 1 is not rejected.
    Like the reference, ANY failure of the root find returns 1.0 - in particular muhat == 0, where 0*log(0) = nan
    (so a top-of-grid risk of exactly 0 makes the sweep stop at its first step; SURVEY.md §7 hard part 1d).
    """
    def _tailprob(mu):
        return min(hoeffding_plus(mu, muhat, n), bentkus_plus(mu, muhat, n)) - np.log(delta)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # nan/0-division warnings are part of the reference's behaviour
        if _tailprob(1 - 1e-10) > 0:
            return 1
        try:
            return brentq(_tailprob, muhat, 1 - 1e-10, maxiter=maxiters)
        except Exception:
            print(f"BRENTQ RUNTIME ERROR at muhat={muhat}")
            return 1.0


def _hb_quiet(muhat, n, delta):
    """HB_mu_plus without the reference's diagnostic print (used only for the bracket search below)."""
    def _tailprob(mu):
        return min(hoeffding_plus(mu, muhat, n), bentkus_plus(mu, muhat, n)) - np.log(delta)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if _tailprob(1 - 1e-10) > 0:
            return 1
        try:
            return brentq(_tailprob, muhat, 1 - 1e-10, maxiter=1000)
        except Exception:
            return 1.0


@functools.lru_cache(maxsize=256)
def hb_stop_bracket(n: int, alpha: float, delta: float):
    """(r_lo, r_hi) with HB_mu_plus(r_lo) <= alpha < HB_mu_plus(r_hi) for strictly positive muhat, r_hi - r_lo tiny.

    HB_mu_plus is non-decreasing in muhat for muhat > 0 (both tail bounds are non-decreasing in the observed mean),
    so ``HB_mu_plus(m) > alpha`` holds for m > r_hi and fails for m < r_lo, up to brentq's 2e-12 tolerance - callers
    add their own guard band and fall back to the real HB_mu_plus inside it.  Special values:
      (0.0, 0.0)  every positive muhat already exceeds alpha (small n);  (inf, inf)  no muhat in (0,1] does.
    muhat == 0 is NOT covered (the reference returns 1.0 there); callers must treat it separately.
    """
    def exceeds(m):
        return _hb_quiet(m, n, delta) > alpha

    tiny = 1e-12
    if exceeds(tiny):
        return 0.0, 0.0
    hi = 1.0
    if not exceeds(hi):
        return math.inf, math.inf
    lo = tiny
    for _ in range(80):
        if hi - lo <= 1e-13 * max(hi, 1e-3):
            break
        mid = 0.5 * (lo + hi)
        if exceeds(mid):
            hi = mid
        else:
            lo = mid
    return lo, hi
