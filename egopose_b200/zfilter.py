"""utils/zfilter.py mirror (class path kept importable through egopose_b200/compat/utils/zfilter.py so the
reference's pickled checkpoints, which store the ZFilter instance, load: ego_pose/ego_mimic.py:133-139).

Sequential semantics (__call__) are the reference's.  The fused rollout freezes (mean, std) for one
rollout and merges the visited observations afterwards with ``merge_moments`` (Chan et al.), a documented
deviation from the reference's sequential worker-0 updates (SURVEY.md 7 'ZFilter semantics')."""
import numpy as np


class RunningStat(object):
    def __init__(self, shape):
        self._n = 0
        self._M = np.zeros(shape)
        self._S = np.zeros(shape)

    def push(self, x):
        x = np.asarray(x)
        assert x.shape == self._M.shape
        self._n += 1
        if self._n == 1:
            self._M[...] = x
        else:
            old = self._M.copy()
            self._M[...] = old + (x - old) / self._n
            self._S[...] = self._S + (x - old) * (x - self._M)

    @property
    def n(self):
        return self._n

    @property
    def mean(self):
        return self._M

    @property
    def var(self):
        return self._S / (self._n - 1) if self._n > 1 else np.square(self._M)

    @property
    def std(self):
        return np.sqrt(self.var)

    @property
    def shape(self):
        return self._M.shape

    def merge_moments(self, n_b, sum_shifted, sumsq_shifted, shift):
        """fold a batch given as sum(x - shift), sum((x - shift)^2) over n_b rows (egp_col_moments_f64)"""
        if n_b <= 0:
            return
        d_mean = sum_shifted / n_b
        mean_b = shift + d_mean
        S_b = np.maximum(sumsq_shifted - n_b * d_mean * d_mean, 0.0)
        if self._n == 0:
            self._n, self._M, self._S = int(n_b), mean_b.copy(), S_b.copy()
            return
        n = self._n + n_b
        delta = mean_b - self._M
        self._S = self._S + S_b + delta * delta * (self._n * n_b / n)
        self._M = self._M + delta * (n_b / n)
        self._n = int(n)


class ZFilter:
    """y = (x - mean) / (std + 1e-8), clipped"""

    def __init__(self, shape, demean=True, destd=True, clip=10.0):
        self.demean = demean
        self.destd = destd
        self.clip = clip
        self.rs = RunningStat(shape)

    def __call__(self, x, update=True):
        if update:
            self.rs.push(x)
        if self.demean:
            x = x - self.rs.mean
        if self.destd:
            x = x / (self.rs.std + 1e-8)
        if self.clip:
            x = np.clip(x, -self.clip, self.clip)
        return x

    def set_mean_std(self, mean, std, n):
        self.rs._n = n
        self.rs._M[...] = mean
        self.rs._S[...] = std
