"""Analytic initial shapes used by the Vortex experiments (reference interface:
core/ana_profiles.py -- vortex(xr, yr, Lx, Ly, x0, y0, sigma, vortex_type, ratio=1))."""
import numpy as np

# radial profiles f(d, sigma), d = (elliptic) distance to the centre; the expressions keep
# the reference's operation order, so the fields are bit-identical to its
SHAPES = {
    'gaussian': lambda d, sigma: np.exp(-d**2/(sigma**2)),
    'cosine': lambda d, sigma: np.where(d > sigma, 0., np.cos(d/sigma*np.pi/2)),
    'step': lambda d, sigma: np.where(d <= sigma, 1., 0.),
}


def vortex(xr, yr, Lx, Ly, x0, y0, sigma, vortex_type, ratio=1):
    """unit-amplitude bump centred at (x0*Lx, y0*Ly); ratio != 1 stretches it in y"""
    if vortex_type not in SHAPES:
        print('this kind of vortex (%s) is not defined' % vortex_type)
        return np.zeros_like(xr*1.)
    distance = np.sqrt((xr-Lx*x0)**2+(yr-Ly*y0)**2*ratio**2)
    return SHAPES[vortex_type](distance, sigma)
