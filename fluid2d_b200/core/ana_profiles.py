"""analytic initial shapes used by the Vortex experiments (reference: core/ana_profiles.py)"""
import numpy as np


def vortex(xr, yr, Lx, Ly, x0, y0, sigma, vortex_type, ratio=1):
    d = np.sqrt((xr-Lx*x0)**2+(yr-Ly*y0)**2*ratio**2)
    y = d*0.
    if vortex_type == 'gaussian':
        y = np.exp(-d**2/(sigma**2))
    elif vortex_type == 'cosine':
        y = np.cos(d/sigma*np.pi/2)
        y[d > sigma] = 0.
    elif vortex_type == 'step':
        y[d <= sigma] = 1.
    else:
        print('this kind of vortex (%s) is not defined' % vortex_type)
    return y
