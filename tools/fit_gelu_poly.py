"""Fit the GELU evaluation used by the GEMM epilogues (common.cuh: gelu_erf / dgelu_erf).

Exact-erf GELU (timm Mlp -> nn.GELU(approximate='none')):  gelu(x) = x * Phi(x).
The epilogue evaluates   Phi(x) ~= sigmoid(x * P(x^2)) = 1 / (1 + exp2(x * Q(x^2)))   with Q = -log2(e) * P, so that
one element costs 2 MUFU (ex2, rcp) + ~8 FMA-pipe instructions; t = x^2 is clamped to XMAX^2 so the fitted
polynomial is never evaluated outside its range (beyond it the sigmoid is saturated to 1e-9).
P is an even polynomial fitted to logit(Phi(x)) / x on |x| <= XMAX by least squares weighted towards the minimax
error of x * Phi(x); the derivative used in backward is the exact derivative of the approximation:
    gelu'(x) = s + x * s * (1 - s) * (P(t) + 2 t P'(t)),   s = sigmoid(x P(t)), t = x^2.
Prints the coefficients (Q, and R = P + 2 t P') and the measured max errors in fp32 arithmetic.
"""
import numpy as np
from scipy.special import erf, ndtr, log_ndtr

XMAX = 5.5


def fit(ncoef, iters=200):
    x = np.linspace(1e-3, XMAX, 6001)
    t = x * x
    # logit(Phi(x)) computed stably: log Phi(x) - log Phi(-x)
    logit = log_ndtr(x) - log_ndtr(-x)
    target = logit / x
    A = np.stack([t ** k for k in range(ncoef)], 1)
    w = np.ones_like(x)
    best = None
    for _ in range(iters):
        # d gelu / d P = x * s(1-s) * x
        s = ndtr(x)
        sens = x * x * s * (1 - s)
        W = w * sens
        coef, *_ = np.linalg.lstsq(A * W[:, None], target * W, rcond=None)
        err = np.abs(x / (1 + np.exp(-x * (A @ coef))) - x * s)
        if best is None or err.max() < best[1]:
            best = (coef.copy(), err.max())
        w = w * (1 + 4 * err / err.max())
        w /= w.mean()
    return best[0]


def evaluate(P):
    P32 = P.astype(np.float32)
    Q32 = (-np.log2(np.e) * P).astype(np.float32)
    R = np.array([(2 * k + 1) * P[k] for k in range(len(P))])  # P + 2 t P'
    R32 = R.astype(np.float32)
    x = np.linspace(-12, 12, 480001).astype(np.float32)
    t = np.minimum(x * x, np.float32(XMAX * XMAX))
    q = np.full_like(x, Q32[-1])
    r = np.full_like(x, R32[-1])
    for k in range(len(P) - 2, -1, -1):
        q = q * t + Q32[k]
        r = r * t + R32[k]
    with np.errstate(over="ignore"):
        s = np.float32(1) / (np.float32(1) + np.exp2(x * q))
    g = x * s
    dg = s + x * s * (np.float32(1) - s) * r
    xd = x.astype(np.float64)
    ref = xd * ndtr(xd)
    dref = ndtr(xd) + xd * np.exp(-0.5 * xd * xd) / np.sqrt(2 * np.pi)
    return Q32, R32, np.abs(g - ref).max(), np.abs(dg - dref).max()


if __name__ == "__main__":
    for n in (2, 3, 4):
        P = fit(n)
        Q, R, e, de = evaluate(P)
        print(f"{n} coefficients: max |gelu err| = {e:.2e}, max |gelu' err| = {de:.2e}")
        print("   Q (exp2 argument / x):", ", ".join(f"{c:.9e}f" for c in Q))
        print("   R (P + 2tP')         :", ", ".join(f"{c:.9e}f" for c in R))
