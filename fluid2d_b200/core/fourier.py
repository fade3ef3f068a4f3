"""Spectral inversion of the SQG model on the device: psi_hat = -pv_hat/|k|, moved half a cell
to the corners, and the diagnosed vorticity |k| pv_hat (reference interface: core/fourier.py,
`Fourier(param, grid).invert(pv, psi, vor)`).

The transforms are a library FFT (cuFFT through torch.fft, complex128) where the reference
calls numpy.fft on the host; the fields never leave HBM and nothing here is a hand-written
kernel.  The two spectral multipliers are tabulated once on the host, with the arithmetic of
fourier.py:19-31 so that they hold the same doubles, and uploaded."""
import numpy as np
import torch


def cell_centres_and_wavenumbers(n, length):
    """x_i = (i + 1/2) L/n and the angular wavenumbers 2 pi m/L in FFT order (m = 0..n/2-1, -n/2..-1)"""
    m = np.rint(np.fft.fftfreq(n)*n)
    return (np.arange(n)+0.5)*length/n, 2*np.pi*m/length


def multipliers(nx, ny, Lx, Ly, dx, dy):
    """(pv -> psi at corners, pv -> vorticity) on the [ny, nx] spectral grid; the mean mode of
    psi is set to zero"""
    kx = cell_centres_and_wavenumbers(nx, Lx)[1]
    ky = cell_centres_and_wavenumbers(ny, Ly)[1]
    KX, KY = np.meshgrid(kx, ky)
    kmod = np.sqrt(KX**2+KY**2)
    safe = kmod.copy()
    safe[0, 0] = 1.                                  # no division by zero for the mean
    to_corner = np.exp(1j*(KX*dx*0.5+KY*dy*0.5))     # half-cell shift: pv at centres, psi at corners
    to_psi = -(1/safe)*to_corner
    to_psi[0, 0] = 0.
    return to_psi, kmod


class Fourier(object):
    def __init__(self, param, grid, device):
        self.nh = param.nh
        self.nx, self.ny = param.nx, param.ny
        self.x, self.kx = cell_centres_and_wavenumbers(param.nx, param.Lx)
        self.y, self.ky = cell_centres_and_wavenumbers(param.ny, param.Ly)
        self.pv2psi, self.pv2vor = multipliers(param.nx, param.ny, param.Lx, param.Ly, grid.dx, grid.dy)
        self.ktot = self.pv2vor
        self.d_pv2psi = torch.from_numpy(np.ascontiguousarray(self.pv2psi)).to(device)
        self.d_pv2vor = torch.from_numpy(np.ascontiguousarray(self.pv2vor)).to(device)

    def invert(self, pv, psi, vor):
        """pv, psi, vor: device tensors [nyl, nxl]; the interiors of psi and vor are overwritten
        (their halos are the caller's fill_halo, operators.py:411-412)"""
        h = self.nh
        inner = (slice(h, -h), slice(h, -h))
        spectrum = torch.fft.fft2(pv[inner])
        psi[inner] = torch.fft.ifft2(spectrum*self.d_pv2psi).real
        vor[inner] = torch.fft.ifft2(spectrum*self.d_pv2vor).real
