"""Host-side set-up helpers (one-off parameter sampling; not on the per-step path).

These follow the semantics of the reference's helpers so that populations built
with the same params and NumPy seed get the same parameters:
``utils.rotate`` (ratinabox/utils.py:293-301), ``utils.distribution_sampler``
(:460-538) and ``utils.create_random_assembly`` (:1115-1219).
"""
import numpy as np


def rotate(vector, theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.matmul(np.array([[c, -s], [s, c]]), vector)


def distribution_sampler(distribution_name="uniform", distribution_parameters=(1,), shape=(10,)):
    prm = distribution_parameters
    prm = tuple(prm) if isinstance(prm, (list, tuple)) else (prm,)
    if distribution_name == "uniform":
        low, high = (0.5 * prm[0], 1.5 * prm[0]) if len(prm) == 1 else (prm[0], prm[1])
        return np.random.uniform(low, high, size=shape)
    if distribution_name == "rayleigh":
        return np.random.rayleigh(scale=prm[0], size=shape)
    if distribution_name == "normal":
        return np.random.normal(loc=prm[0], scale=prm[1], size=shape)
    if distribution_name == "logarithmic":
        assert len(shape) == 1, "Logarithmic distribution only works for 1D arrays"
        return np.logspace(np.log10(prm[0]), np.log10(prm[1]), num=shape[0], base=10)
    if distribution_name == "delta":
        return prm[0] * np.ones(shape)
    if distribution_name == "modules":
        assert len(shape) == 1, "Modules distribution only works for 1D arrays"
        per = shape[0] // len(prm)
        out = prm[-1] * np.ones(shape)          # remainder goes to the last module
        for i, val in enumerate(prm):
            out[i * per:(i + 1) * per] = val
        return out
    if distribution_name == "truncnorm":
        import scipy.stats
        lower, upper, mu, sigma = prm[:4]
        return scipy.stats.truncnorm.rvs((lower - mu) / sigma, (upper - mu) / sigma, scale=sigma, loc=mu, size=shape)
    raise ValueError("This distribution is not recognised")


def create_random_assembly(tuning_distance_distribution="uniform", tuning_distance=(0.02, 0.3),
                           tuning_angle_distribution="uniform", tuning_angle=(0.0, 360.0),
                           sigma_angle_distribution="uniform", sigma_angle=(10, 30),
                           sigma_distance_distribution="diverging", sigma_distance=(0.08, 12), n=10, **kwargs):
    """Random vector-cell tuning: (tuning_distance, tuning_angle [rad], sigma_distance, sigma_angle [rad]).
    Draw order (distance, [sigma_d], angle, sigma_angle) matches the reference so seeds line up."""
    given = [p for p in (tuning_distance, tuning_angle, sigma_distance, sigma_angle) if type(p) in (list, np.ndarray)]
    if given:
        lengths = {len(p) for p in given}
        assert len(lengths) == 1, "If more than one parameter is passed as a list, they must all have the same length"
        n = lengths.pop()

    def draw(value, dist):
        if type(value) in (list, np.ndarray):
            return np.array(value, dtype=float)
        return distribution_sampler(dist, value, (n,))

    mu_d = np.abs(draw(tuning_distance, tuning_distance_distribution))
    if type(sigma_distance) in (list, np.ndarray):
        sg_d = np.array(sigma_distance, dtype=float)
    elif sigma_distance_distribution == "diverging":
        sg_d = sigma_distance[0] + mu_d / sigma_distance[1]        # Hartley: xi + mu/beta
    else:
        sg_d = distribution_sampler(sigma_distance_distribution, sigma_distance, (n,))
    mu_t = draw(tuning_angle, tuning_angle_distribution) * (np.pi / 180)
    sg_t = draw(sigma_angle, sigma_angle_distribution) * (np.pi / 180)
    return mu_d, mu_t, sg_d, sg_t


def create_uniform_radial_assembly(distance_range=(0.0, 0.2), angle_range=(0, 90), spatial_resolution=0.04, **kwargs):
    """Concentric rows of equally sized receptive fields tiling a field of view (the reference's
    utils.create_uniform_radial_assembly, ratinabox/utils.py:1033-1070)."""
    lo, hi = (a * np.pi / 180 for a in angle_range)
    mu_d, mu_t, sg_d, sg_t = [], [], [], []
    for radius in np.arange(max(0.01, distance_range[0]), distance_range[1], spatial_resolution):
        dtheta = spatial_resolution / radius
        right = np.arange(lo + dtheta / 2, hi, dtheta)
        for theta in np.concatenate((-right[::-1], right)):
            mu_d.append(radius); mu_t.append(theta); sg_d.append(spatial_resolution); sg_t.append(spatial_resolution / radius)
    return mu_d, mu_t, sg_d, sg_t


def create_diverging_radial_assembly(distance_range=(0.01, 0.2), angle_range=(0, 90), spatial_resolution=0.04,
                                     beta=5, **kwargs):
    """Rows whose receptive fields grow with radius (Hartley et al. 2000): sigma_d = xi + radius/beta with xi fixed
    by the innermost row (the reference's utils.create_diverging_radial_assembly, ratinabox/utils.py:1073-1112)."""
    lo, hi = (a * np.pi / 180 for a in angle_range)
    mu_d, mu_t, sg_d, sg_t = [], [], [], []
    radius = max(0.01, distance_range[0])
    xi = spatial_resolution - radius / beta
    while radius < distance_range[1]:
        res = xi + radius / beta
        dtheta = res / radius
        right = np.array([lo + dtheta / 2]) if dtheta / 2 > hi else np.arange(lo + dtheta / 2, hi, dtheta)
        for theta in np.concatenate((-right[::-1], right)):
            mu_d.append(radius); mu_t.append(theta); sg_d.append(res); sg_t.append(res / radius)
        radius = (2 * radius + res + xi) / (2 - 1 / beta)      # next row just touches this one
    return mu_d, mu_t, sg_d, sg_t


def bin_data_for_histogramming(data, extent, dx, weights=None, norm_by_bincount=False, return_zero_bins=False):
    """ratinabox.utils.bin_data_for_histogramming, 2D branch (utils.py:544-589): np.histogram2d over
    np.arange(extent[0], extent[1] + dx, dx) x np.arange(extent[2], extent[3] + dx, dx), optionally weighted and
    divided by the bin count, returned `.T[::-1, :]`.  Host NumPy; the device path over the history rings is
    ``Agent.get_position_heatmap`` / ``Neurons.get_history_rate_maps``."""
    assert len(extent) == 4, "2D only"
    data = np.asarray(data, dtype=float)
    bins_x = np.arange(extent[0], extent[1] + dx, dx)
    bins_y = np.arange(extent[2], extent[3] + dx, dx)
    heatmap, _, _ = np.histogram2d(data[:, 0], data[:, 1], bins=[bins_x, bins_y], weights=weights)
    zero_bins = None
    if norm_by_bincount:
        bincount, _, _ = np.histogram2d(data[:, 0], data[:, 1], bins=[bins_x, bins_y])
        zero_bins = (bincount == 0)
        bincount[zero_bins] = 1
        heatmap = heatmap / bincount
    heatmap = heatmap.T[::-1, :]
    if return_zero_bins:
        return (heatmap, zero_bins.T[::-1, :])
    return heatmap
