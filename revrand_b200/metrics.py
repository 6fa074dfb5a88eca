"""Regression scores used to report results next to the reference
(revrand/metrics.py:9-141): host-side, O(N)."""

import numpy as np
from scipy.stats import norm


def smse(y_true, y_pred):
    """Mean squared error standardised by the variance of the targets."""
    y_true, y_pred = np.asarray(y_true), np.asarray(y_pred)
    return np.mean((y_true - y_pred) ** 2) / np.var(y_true)


def mll(y_true, y_pred, y_var):
    """Mean log LOSS under a Gaussian predictive distribution: the mean
    NEGATIVE log-likelihood, lower is better (revrand/metrics.py:38-66)."""
    return -np.mean(norm.logpdf(y_true, loc=y_pred, scale=np.sqrt(y_var)))


def msll(y_true, y_pred, y_var, y_train):
    """Mean standardised log loss against the trivial Gaussian predictor
    fitted to the training targets (negative is better)."""
    base = norm.logpdf(y_true, loc=np.mean(y_train), scale=np.std(y_train))
    return -(np.mean(norm.logpdf(y_true, loc=y_pred,
                                 scale=np.sqrt(y_var))) - np.mean(base))


def lins_ccc(y_true, y_pred):
    """Lin's concordance correlation coefficient."""
    y_true, y_pred = np.asarray(y_true), np.asarray(y_pred)
    cov = np.mean((y_true - y_true.mean()) * (y_pred - y_pred.mean()))
    return 2 * cov / (y_true.var() + y_pred.var()
                      + (y_true.mean() - y_pred.mean()) ** 2)
