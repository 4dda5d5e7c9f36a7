"""Likelihood nodes of a final DGP layer (dgpsi/likelihood_class.py): Poisson (:8-90), Hetero (:92-243) and NegBin
(:245-292) -- SURVEY.md section 8f-3.

The log-likelihood that enters the ESS acceptance rule is evaluated on the device (`dgpb_lik_loglik`; inside an
I-step by `dgpb_ess_block_lik`, which keeps every proposal in HBM).  `Hetero.posterior` draws the mean process from
its exact Gaussian conditional with the sliding-window factorisation (`dgpb_mvn_draw`,
`dgpb_compute_stats_shifted`).  `prediction` / `sampling` / `pllik` are elementwise host formulas on the moments the
prediction kernels return.  `Categorical` (:294-468): two classes through a logit / probit link, K classes through
softmax / robustmax.  `ZIP` (:470-622) and `ZINB` (:624-814): zero-inflated Poisson / negative-binomial counts.
Out of scope: replicate pooling (`rep`), Hetero under Vecchia.
"""
from __future__ import annotations

import ctypes

import numpy as np
from scipy.special import expit, gammaln, log_ndtr, ndtr, owens_t

from . import _lib as L


class _Likelihood:
    n_inputs = 1
    exact_post_idx = None

    def __init__(self, input_dim=None):
        self.type = 'likelihood'
        self.name = type(self).__name__
        self.input = None
        self.output = None
        self.input_dim = input_dim
        self.exact_post_idx = type(self).exact_post_idx
        self.rep = None

    def _kind(self):
        return L.LIK_KIND[self.name]

    def _descriptor(self, rows, y_dev):
        d = L.DgpbLik()
        d.kind = self._kind()
        rows = [int(r) for r in rows]
        if len(rows) > 8:
            raise NotImplementedError("dgp_b200: a likelihood node reads at most 8 latent columns")
        d.n_in = len(rows)
        for j, r in enumerate(rows):
            d.rows[j] = r
        d.y = y_dev.data_ptr()
        d.param = float(getattr(self, 'robustmax_eps', 0.0))
        return d

    def llik(self):
        """Sum of the log-likelihood of the observed outputs given the latent inputs, evaluated on the device."""
        F = L.to_dev(np.ascontiguousarray(self.input.T))
        y = L.to_dev(np.ascontiguousarray(self.output[:, 0], dtype=np.float64))
        d = (L.DgpbLik * 1)(self._descriptor(range(self.input.shape[1]), y))
        out = ctypes.c_double(0.0)
        L.check(L.load().dgpb_lik_loglik(d, 1, L.ptr(F), F.shape[1], ctypes.byref(out), L.stream()))
        return out.value


class Poisson(_Likelihood):
    """Poisson likelihood fed by one GP node, the log rate (likelihood_class.py:8-90)."""
    n_inputs = 1

    @staticmethod
    def pllik(y, f):
        return y * f - np.exp(f) - gammaln(y + 1)

    @staticmethod
    def prediction(m, v):
        y_mean = np.exp(m + v / 2)
        y_var = np.exp(m + v / 2) + (np.exp(v) - 1) * np.exp(2 * m + v)
        return y_mean.flatten(), y_var.flatten()

    def sampling(self, f_sample):
        return np.random.poisson(np.exp(f_sample)).flatten()


class Hetero(_Likelihood):
    """Heteroskedastic Gaussian likelihood fed by two GP nodes: mean and log variance (likelihood_class.py:92-243).
    The mean node has a closed-form conditional posterior (`exact_post_idx = [0]`)."""
    n_inputs = 2
    exact_post_idx = np.array([0])

    @staticmethod
    def pllik(y, f):
        mu, var = f[:, :, [0]], np.exp(f[:, :, [1]])
        return -0.5 * (np.log(2 * np.pi * var) + (y - mu) ** 2 / var)

    @staticmethod
    def prediction(m, v):
        y_mean = m[:, 0]
        y_var = np.exp(m[:, 1] + v[:, 1] / 2) + v[:, 0]
        return y_mean.flatten(), y_var.flatten()

    @staticmethod
    def sampling(f_sample):
        return np.random.normal(f_sample[:, 0], np.sqrt(np.exp(f_sample[:, 1]))).flatten()

    @staticmethod
    def posterior_dev(node, n, log_var, y, sd):
        """Draw of the mean process f | y, log variance (post_het1, likelihood_class.py:185-210) on the device.
        With v = scale K the prior covariance and G = diag(exp(log_var)): u = chol(v) sd0, w = sqrt(G) sd1,
        f = mu + u - v (v+G)^-1 (u+w), mu = v (v+G)^-1 y.  Writing x = (v+G)^-1 (y-u-w) this is u + v x, and since
        v x = (y-u-w) - G x the draw is f = y - w - G x: one shifted factorisation, no matrix-vector product.
        node: DgpbNode of the mean GP; log_var, y: device vectors; sd: (n x 2) standard normals (host)."""
        torch = L.torch_mod()
        lib = L.load()
        z0, z1 = L.to_dev(np.ascontiguousarray(sd[:, 0])), L.to_dev(np.ascontiguousarray(sd[:, 1]))
        u = L.empty((n,))
        L.check(lib.dgpb_mvn_draw(L.workspace(), ctypes.byref(node), n, L.ptr(z0), L.ptr(u), L.stream()))
        gamma = torch.exp(log_var)
        w = torch.sqrt(gamma) * z1
        rhs = (y - u - w).contiguous()
        shift = (gamma / node.scale).contiguous()       # v + G = scale (K + G / scale)
        tmp = L.DgpbNode.from_buffer_copy(node)
        tmp.output = rhs.data_ptr()
        Rinv, x = L.empty((n, n)), L.empty((n,))
        L.check(lib.dgpb_compute_stats_shifted(L.workspace(), ctypes.byref(tmp), n, L.ptr(shift), L.ptr(Rinv), L.ptr(x),
                                               L.stream()))
        return y - w - shift * x                          # G x_true = (G / scale) x

    @staticmethod
    def posterior_vecch_dev(kern, n, log_var, y, sd):
        """Draw of the mean process under the Vecchia approximation (imputation.py:141-158, U_matrix_sp,
        post_het_vecch): `dgpb_hetero_vecchia_draw` on the inputs, outputs and variances gathered in Vecchia order.
        kern: the GP node of the mean (a `dgp_b200.kernel` with `imp_NNarray`); log_var, y: device vectors in data
        order; sd: n standard normals (host).  Returns the draw in data order (device)."""
        torch = L.torch_mod()
        ordd = L.to_dev(kern.ord, np.int64)
        Xo = L.to_dev(np.ascontiguousarray(kern._X()[kern.ord]))
        NN = L.to_dev(np.ascontiguousarray(kern.imp_NNarray), np.int64)
        gamma = torch.exp(log_var).index_select(0, ordd).contiguous()
        yo = y.index_select(0, ordd).contiguous()
        z = L.to_dev(np.ascontiguousarray(sd, dtype=np.float64))
        f = L.empty((n,))
        larr, lptr = L.length_host(kern.length)
        L.check(L.load().dgpb_hetero_vecchia_draw(L.ptr(Xo), L.ptr(NN), n, Xo.shape[1], NN.shape[1], lptr, len(larr),
                                                  float(kern.scale[0]), L.KIND[kern.name], L.ptr(gamma), L.ptr(yo),
                                                  L.ptr(z), L.ptr(f), None, L.stream()))
        out = torch.empty_like(f)
        out.index_copy_(0, ordd, f)     # f[rev_ord]
        return out

    def posterior(self, idx, v_node):
        """Host-facing form of `Hetero.posterior` (likelihood_class.py:134-151): `v_node` is the GP node (a
        `dgp_b200.kernel`) that produces the mean; returns the drawn mean vector."""
        if int(np.atleast_1d(idx)[0]) != 0:
            return None
        if self.rep is not None:
            raise NotImplementedError("dgp_b200: replicate pooling is outside the SI hot path")
        n = len(self.output)
        bufs = v_node._upload()
        node = v_node._node(bufs)
        sd = np.random.randn(n, 2)                        # likelihood_class.py:200
        f = self.posterior_dev(node, n, L.to_dev(np.ascontiguousarray(self.input[:, 1])),
                               L.to_dev(np.ascontiguousarray(self.output[:, 0])), sd)
        return L.to_host(f)


class NegBin(_Likelihood):
    """Negative-binomial likelihood fed by two GP nodes: log mean and log dispersion (likelihood_class.py:245-292)."""
    n_inputs = 2

    @staticmethod
    def pllik(y, f):
        f1, f2 = f[:, :, [0]], f[:, :, [1]]
        n = np.exp(-f2)
        a = f1 + f2
        return gammaln(y + n) - gammaln(n) - gammaln(y + 1.0) + y * a - (y + n) * np.logaddexp(0.0, a)

    @staticmethod
    def prediction(m, v):
        y_mean = np.exp(m[:, 0] + v[:, 0] / 2)
        y_var = (np.exp(2 * m[:, 0] + v[:, 0]) * (np.exp(v[:, 0]) - 1) + np.exp(m[:, 0] + v[:, 0] / 2)
                 + np.exp(m[:, 1] + v[:, 1] / 2) * np.exp(2 * m[:, 0] + 2 * v[:, 0]))
        return y_mean.flatten(), y_var.flatten()

    @staticmethod
    def sampling(f_sample):
        p, k = 1 / (1 + np.exp(f_sample[:, 0] + f_sample[:, 1])), np.exp(-f_sample[:, 1])
        return np.random.negative_binomial(k, p).flatten()


class Categorical(_Likelihood):
    """Categorical likelihood (likelihood_class.py:294-468): `num_classes == 2` reads one GP node through a
    'logit' or 'probit' link, K > 2 classes read K GP nodes through 'softmax' or 'robustmax'.  `dgp.__init__`
    encodes the labels (`class_encoder`) and fills `num_classes` / `link` when they are None."""

    def __init__(self, num_classes=None, input_dim=None, link=None, robustmax_eps=1e-3):
        super().__init__(input_dim)
        self.num_classes = num_classes
        self.class_encoder = None
        self.link = link
        self.robustmax_eps = robustmax_eps

    @property
    def n_inputs(self):
        return 1 if self.num_classes == 2 else self.num_classes

    def _kind(self):
        if self.num_classes == 2:
            return L.CAT_KIND['logit' if self.link == 'logit' else 'probit']
        return L.CAT_KIND['robustmax' if self.link == 'robustmax' else 'softmax']

    def pllik(self, y, f):
        if self.num_classes == 2:
            if self.link == 'logit':
                return y * f - np.logaddexp(0, f)
            return y * log_ndtr(f) + (1 - y) * log_ndtr(-f)
        labels = y.flatten().astype(int)
        if self.link == 'robustmax':
            hit = np.argmax(f, axis=2) == labels[:, None]
            return np.where(hit, np.log(1.0 - self.robustmax_eps),
                            np.log(self.robustmax_eps / (self.num_classes - 1)))[:, :, None]
        top = np.max(f, axis=2, keepdims=True)
        lse = np.log(np.sum(np.exp(f - top), axis=2)) + np.squeeze(top, axis=2)
        return (f[np.arange(len(labels)), :, labels] - lse)[:, :, None]

    def prediction(self, m, v):
        """Class probabilities (mean) and their variances from Gaussian moments of the latent inputs
        (likelihood_class.py:384-449): closed forms for two classes, 1000 Monte-Carlo draws from numpy's global
        RNG -- in the reference's order and shapes -- for K classes."""
        if self.num_classes == 2:
            m, v = m.flatten(), v.flatten()
            if self.link == 'logit':
                denom = 1.0 + (np.pi / 8.0) * v
                y_mean = expit(m / np.sqrt(denom))
                y_var = np.clip((y_mean * (1.0 - y_mean)) ** 2 * (v / denom), 0.0, y_mean * (1.0 - y_mean))
            else:
                t = m / np.sqrt(1.0 + v)
                y_mean = ndtr(t)
                y_var = np.maximum(y_mean - 2.0 * owens_t(t, 1.0 / np.sqrt(1.0 + 2.0 * v)) - y_mean * y_mean, 0.0)
            return y_mean.reshape(-1, 1), y_var.reshape(-1, 1)
        K, S, chunk = self.num_classes, 1000, 200
        std = np.sqrt(np.maximum(v, 0.0))
        M = m.shape[0]
        if self.link == 'robustmax':
            wins = np.zeros((M, K))
            for _ in range(S // chunk):
                f = m[:, None, :] + std[:, None, :] * np.random.randn(M, chunk, K)
                np.add.at(wins, (np.arange(M)[:, None], np.argmax(f, axis=2)), 1.0)
            q = wins / S
            a, b = 1.0 - self.robustmax_eps, self.robustmax_eps / (K - 1)
            return b + (a - b) * q, (a - b) ** 2 * q * (1.0 - q)
        sum_p, sum_p2 = np.zeros((M, K)), np.zeros((M, K))
        for _ in range(S // chunk):
            half = np.random.randn(M, (chunk + 1) // 2, K)             # antithetic pairs
            f = m[:, None, :] + std[:, None, :] * np.concatenate([half, -half], axis=1)[:, :chunk, :]
            f -= np.max(f, axis=2, keepdims=True)
            p = np.exp(f)
            p /= np.sum(p, axis=2, keepdims=True)
            sum_p += p.sum(axis=1)
            sum_p2 += (p * p).sum(axis=1)
        y_mean = sum_p / S
        return y_mean, sum_p2 / S - y_mean ** 2

    def sampling(self, f_sample):
        if self.num_classes == 2:
            return expit(f_sample) if self.link == 'logit' else ndtr(f_sample)
        if self.link == 'robustmax':
            out = np.full_like(f_sample, self.robustmax_eps / (self.num_classes - 1), dtype=float)
            out[np.arange(f_sample.shape[0]), np.argmax(f_sample, axis=1)] = 1.0 - self.robustmax_eps
            return out
        e = np.exp(f_sample - np.max(f_sample, axis=1, keepdims=True))
        return e / np.sum(e, axis=1, keepdims=True)


def _logit_normal_moments(m_pi, v_pi):
    """Mean and variance of expit(g), g ~ N(m_pi, v_pi), by the probit-style approximation used by the reference
    (likelihood_class.py:585-592, 766-773)."""
    denom = np.maximum(1.0 + (np.pi / 8.0) * v_pi, 1e-12)
    pi_mean = expit(m_pi / np.sqrt(denom))
    pi_var = np.clip((pi_mean * (1.0 - pi_mean)) ** 2 * (v_pi / denom), 0.0, pi_mean * (1.0 - pi_mean))
    return pi_mean, pi_var


def _zero_inflate(log_count, f_pi, is_zero):
    """log[(1 - pi) p(y) + pi 1(y = 0)] with pi = expit(f_pi) and log p(y) = log_count."""
    pi = expit(f_pi)
    with np.errstate(divide='ignore'):
        inflated = np.logaddexp(np.log(pi), np.log1p(-pi) + log_count)
    return np.where(is_zero, inflated, np.log1p(-pi) + log_count)


class ZIP(_Likelihood):
    """Zero-inflated Poisson fed by two GP nodes: log rate and logit of the zero-inflation probability
    (likelihood_class.py:470-622)."""
    n_inputs = 2

    @staticmethod
    def pllik(y, f):
        f_lam, f_pi = f[..., 0:1], f[..., 1:2]
        yb = np.broadcast_to(y, f_lam.shape)
        return _zero_inflate(-np.exp(f_lam) + yb * f_lam - gammaln(yb + 1.0), f_pi, yb == 0)

    @staticmethod
    def prediction(m, v):
        lam_mean = np.exp(m[:, 0] + 0.5 * v[:, 0])
        lam_var = (np.exp(v[:, 0]) - 1.0) * np.exp(2.0 * m[:, 0] + v[:, 0])
        pi_mean, pi_var = _logit_normal_moments(m[:, 1], v[:, 1])
        y_mean = (1.0 - pi_mean) * lam_mean
        cond_var = (1.0 - pi_mean) * lam_mean * (1.0 + pi_mean * lam_mean)
        var_g = ((1.0 - pi_mean) ** 2 + pi_var) * lam_var + pi_var * lam_mean ** 2
        return y_mean.flatten(), np.maximum(cond_var + var_g, 0.0).flatten()

    def sampling(self, f_sample):
        lam, pi = np.exp(f_sample[:, 0]), expit(f_sample[:, 1])
        u = np.random.rand(f_sample.shape[0])
        return np.where(u < pi, 0, np.random.poisson(lam)).flatten()


class ZINB(_Likelihood):
    """Zero-inflated negative binomial fed by three GP nodes: log mean, log dispersion and logit of the
    zero-inflation probability (likelihood_class.py:624-814)."""
    n_inputs = 3

    @staticmethod
    def pllik(y, f):
        f1, f2, f_pi = f[..., 0:1], f[..., 1:2], f[..., 2:3]
        n, a = np.exp(-f2), f1 + f2
        yb = np.broadcast_to(np.asarray(y), n.shape)
        log_nb = gammaln(yb + n) - gammaln(n) - gammaln(yb + 1.0) + yb * a - (yb + n) * np.logaddexp(0.0, a)
        return _zero_inflate(log_nb, f_pi, yb == 0)

    @staticmethod
    def prediction(m, v):
        m1, v1, m2, v2 = m[:, 0], v[:, 0], m[:, 1], v[:, 1]
        mu_mean = np.exp(m1 + 0.5 * v1)
        mu_var = (np.exp(v1) - 1.0) * np.exp(2.0 * m1 + v1)
        mu2_mean = np.exp(2.0 * m1 + 2.0 * v1)
        mu2_over_n = mu2_mean * np.exp(m2 + 0.5 * v2)
        pi_mean, pi_var = _logit_normal_moments(m[:, 2], v[:, 2])
        y_mean = (1.0 - pi_mean) * mu_mean
        e_pi1m = np.clip(pi_mean * (1.0 - pi_mean) - pi_var, 0.0, pi_mean * (1.0 - pi_mean))
        cond_var = (1.0 - pi_mean) * (mu_mean + mu2_over_n) + e_pi1m * mu2_mean
        var_g = ((1.0 - pi_mean) ** 2 + pi_var) * mu_var + pi_var * mu_mean ** 2
        return y_mean.flatten(), np.maximum(cond_var + var_g, 0.0).flatten()

    @staticmethod
    def sampling(f_sample):
        k, p = np.exp(-f_sample[:, 1]), 1.0 / (1.0 + np.exp(f_sample[:, 0] + f_sample[:, 1]))
        pi = expit(f_sample[:, 2])
        u = np.random.rand(f_sample.shape[0])
        nb = np.random.negative_binomial(k, p)
        return np.where(u < pi, 0, nb).flatten()
