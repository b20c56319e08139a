"""reference: extensions/earth_movers_distance/emd.py -- outside the B200 hot path ('EMD only if coarse_loss == emd is
ever required').  The class is constructible (metrics/loss.py:17 and metrics/metric.py:9 build one at import) and raises when called."""
from torch import nn


class EarthMoverDistance(nn.Module):
    def forward(self, xyz1, xyz2):
        raise NotImplementedError("EarthMoverDistance is outside the B200 hot path (SURVEY.md 8)")
