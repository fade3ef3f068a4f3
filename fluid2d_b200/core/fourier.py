"""Fourier: spectral inversion of the SQG model, psi = -pv/|k| (reference: core/fourier.py).

The transforms are a library FFT on the device (cuFFT through torch.fft, complex128) where the
reference calls numpy.fft on the host; the spectral multipliers are built exactly as the
reference builds them (fourier.py:19-31) and uploaded once.  Nothing here is a hand-written
kernel; the fields never leave HBM."""
import numpy as np
import torch


def set_x_and_k(n, L):
    k = ((n//2+np.arange(n)) % n) - n//2
    return (np.arange(n)+0.5)*L/n, 2*np.pi*k/L


class Fourier(object):
    def __init__(self, param, grid, device):
        dx, dy = grid.dx, grid.dy
        self.nx, self.ny = param.nx, param.ny
        self.Lx, self.Ly = param.Lx, param.Ly
        self.nh = param.nh
        self.x, self.kx = set_x_and_k(self.nx, self.Lx)
        self.y, self.ky = set_x_and_k(self.ny, self.Ly)
        self.xx, self.yy = np.meshgrid(self.x, self.y)
        self.kxx, self.kyy = np.meshgrid(self.kx, self.ky)
        self.ktot = np.sqrt(self.kxx**2+self.kyy**2)
        self.ktot[0, 0] = 1.          # avoid the division by zero of the mean mode
        # half-cell shift in Fourier space: psi lives on cell corners, pv on cell centres
        shift = np.exp(1j*(self.kxx*dx*0.5+self.kyy*dy*0.5))
        self.pv2psi = -(1/self.ktot)*shift
        self.pv2vor = self.ktot
        self.pv2psi[0, 0] = 0.
        self.ktot[0, 0] = 0.
        self.d_pv2psi = torch.from_numpy(np.ascontiguousarray(self.pv2psi)).to(device)
        self.d_pv2vor = torch.from_numpy(np.ascontiguousarray(self.pv2vor)).to(device)

    def invert(self, pv, psi, vor):
        """pv, psi, vor: device tensors [nyl, nxl]; interiors of psi and vor are overwritten"""
        nh = self.nh
        hpv = torch.fft.fft2(pv[nh:-nh, nh:-nh])
        psi[nh:-nh, nh:-nh] = torch.fft.ifft2(hpv*self.d_pv2psi).real
        vor[nh:-nh, nh:-nh] = torch.fft.ifft2(hpv*self.d_pv2vor).real
