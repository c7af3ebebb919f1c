"""``SequenceModel``: full-trajectory GPODE (mirror of reference ``src/gpode/models.py``)."""
from torch import nn

from ..misc.torch_utils import compute_ts_dense, insert_zero_t0


class SequenceModel(nn.Module):
    def __init__(self, flow, num_observations, x0_distribution, likelihood, ts_dense_scale=1):
        super().__init__()
        self.flow = flow
        self.num_observations = num_observations
        self.x0_distribution = x0_distribution
        self.likelihood = likelihood
        self.ts_dense_scale = ts_dense_scale

    def build_flow(self, x0, ts):
        """Integrate ``x0 (N,D)`` over ``ts (T,)`` refined by ``ts_dense_scale`` and return the states at the original
        time points, ``(N,T,D)`` (reference ``models.py:32-43``)."""
        if self.ts_dense_scale < 2:
            # the reference slices with stride ts_dense_scale-1, which is a zero stride for 1
            raise ValueError("ts_dense_scale must be >= 2")
        dense = compute_ts_dense(ts, self.ts_dense_scale)
        ys = self.flow(x0, dense)
        return ys[:, ::self.ts_dense_scale - 1, :]

    def build_lowerbound_terms(self, ys, ts):
        """-> (mean observation log-likelihood, KL[q(x0)||p(x0)] / num_observations) (reference ``models.py:45-58``)."""
        ts = insert_zero_t0(ts)
        x0_samples = self.x0_distribution.sample(num_samples=1)[0]
        x0_kl = self.x0_distribution.kl()
        xs = self.build_flow(x0_samples, ts)[:, 1:]
        mean_fn = getattr(self.likelihood, "log_prob_mean", None)
        loglik_mean = mean_fn(xs, ys) if mean_fn is not None else self.likelihood.log_prob(xs, ys).mean()
        return loglik_mean, x0_kl.mean() / self.num_observations

    def build_kl(self):
        return self.flow.kl() / self.num_observations

    def forward(self, x0, ts):
        return self.build_flow(x0, ts)

    def forward_sets(self, x0, ts, rng="numpy"):
        """``x0 (n,N,D)`` -> ``(n,N,T,D)``: ``forward`` for n independent GP draws in one launch (prediction only)."""
        if self.ts_dense_scale < 2:
            raise ValueError("ts_dense_scale must be >= 2")
        dense = compute_ts_dense(ts, self.ts_dense_scale)
        return self.flow.forward_sets(x0, dense, rng=rng)[:, :, ::self.ts_dense_scale - 1, :]
