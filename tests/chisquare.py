"""The reference's chi-square goodness-of-fit test for sampling routines (src/libcore/chisquare.cpp:46-260), restated
with numpy: samples are binned over the sphere in (theta, phi), the expected counts come from integrating the claimed
density over every bin (plus any discrete directions), low-frequency cells are pooled, and the P-value is compared with
the Sidak-corrected significance level.  Used by tests/test_chisquare.py the way src/tests/test_chisquare.cpp uses it."""
import numpy as np
from scipy import stats

CHISQR_MIN_EXP_FREQUENCY = 5          # chisquare.h
SIGNIFICANCE_LEVEL = 0.0025           # test_chisquare.cpp:30


def to_spherical(d):
    """toSphericalCoordinates (util.cpp): theta = acos(z), phi = atan2(y, x) in [0, 2 pi)."""
    theta = np.arccos(np.clip(d[..., 2], -1, 1))
    phi = np.arctan2(d[..., 1], d[..., 0])
    phi = np.where(phi < 0, phi + 2 * np.pi, phi)
    return theta, phi


def from_spherical(theta, phi):
    s = np.sin(theta)
    return np.stack([s * np.cos(phi), s * np.sin(phi), np.cos(theta)], axis=-1)


class ChiSquare:
    def __init__(self, theta_bins=10, phi_bins=0, num_tests=1, sample_count=0, quad_nodes=12, quad_cells=4):
        self.tb, self.pb = theta_bins, phi_bins or 2 * theta_bins
        self.num_tests = num_tests
        self.sample_count = sample_count or self.tb * self.pb * 1000
        self.tolerance = self.sample_count * 1e-4
        # composite Gauss-Legendre rule per bin (quad_cells sub-intervals of quad_nodes points per axis): densities with a
        # kink or jump inside a bin (critical angle of a rough dielectric) need the subdivision
        x, w = np.polynomial.legendre.leggauss(quad_nodes)
        self.nodes = np.concatenate([(x + 2 * c + 1) / quad_cells - 1 for c in range(quad_cells)])
        self.weights = np.concatenate([w / quad_cells for _ in range(quad_cells)])

    def fill(self, directions, weights, discrete_mask, pdf_fn):
        """directions [n,3], weights [n] (0 for failed samples), discrete_mask [n]; pdf_fn(dirs, discrete) -> [m]."""
        n = self.sample_count
        assert len(directions) == n
        theta, phi = to_spherical(directions)
        ti = np.clip(np.floor(theta * self.tb / np.pi).astype(int), 0, self.tb - 1)
        pi_ = np.clip(np.floor(phi * self.pb / (2 * np.pi)).astype(int), 0, self.pb - 1)
        self.table = np.zeros(self.tb * self.pb)
        np.add.at(self.table, ti * self.pb + pi_, weights)
        self.ref = np.zeros(self.tb * self.pb)
        disc = directions[(weights > 0) & discrete_mask]
        if len(disc):
            uniq = np.unique(disc, axis=0)
            p = pdf_fn(uniq, True)
            t, ph = to_spherical(uniq)
            idx = np.clip(np.floor(t * self.tb / np.pi).astype(int), 0, self.tb - 1) * self.pb \
                + np.clip(np.floor(ph * self.pb / (2 * np.pi)).astype(int), 0, self.pb - 1)
            np.add.at(self.ref, idx, p * n)
        # integral of pdf(theta, phi) sin(theta) over every bin: tensor Gauss-Legendre rule (the reference uses adaptive cubature)
        dt, dp = np.pi / self.tb, 2 * np.pi / self.pb
        tt = (np.arange(self.tb)[:, None] + 0.5 + 0.5 * self.nodes[None, :]) * dt            # [tb, q]
        pp = (np.arange(self.pb)[:, None] + 0.5 + 0.5 * self.nodes[None, :]) * dp            # [pb, q]
        T = tt[:, None, :, None] + 0 * pp[None, :, None, :]
        P = pp[None, :, None, :] + 0 * tt[:, None, :, None]
        dirs = from_spherical(T.ravel(), P.ravel())
        vals = pdf_fn(dirs, False).reshape(T.shape) * np.sin(T)
        w = self.weights[None, None, :, None] * self.weights[None, None, None, :] * (dt / 2) * (dp / 2)
        integ = (vals * w).sum(axis=(2, 3))
        self.integral = float(integ.sum())
        self.ref += integ.ravel() * n

    def run_test(self, pval_thresh=SIGNIFICANCE_LEVEL):
        order = np.argsort(self.ref, kind="stable")
        pooled_counts = pooled_ref = chsq = 0.0
        pooled_cells = df = 0
        for idx in order:
            if self.ref[idx] == 0:
                if self.table[idx] > self.tolerance:
                    return "reject", 0.0
            elif self.ref[idx] < CHISQR_MIN_EXP_FREQUENCY or (0 < pooled_ref < CHISQR_MIN_EXP_FREQUENCY):
                pooled_counts += self.table[idx]; pooled_ref += self.ref[idx]; pooled_cells += 1
            else:
                diff = self.table[idx] - self.ref[idx]
                chsq += diff * diff / self.ref[idx]; df += 1
        if pooled_cells > 0:
            diff = pooled_counts - pooled_ref
            chsq += diff * diff / pooled_ref; df += 1
        df -= 1
        if df <= 0:
            return "low_dof", 1.0
        pval = float(stats.chi2.sf(chsq, df))
        alpha = 1 - (1 - pval_thresh) ** (1.0 / self.num_tests)       # Sidak correction
        return ("reject" if pval < alpha else "accept"), pval
