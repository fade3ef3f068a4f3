"""Experiment set-ups shared by the golden-vector generator (which runs them on the
REFERENCE's Python, tests/golden/make_golden.py) and by the parity tests (which run
them on the product, fluid2d_b200, and on the oracle's restated driver).

Each case is written the way a Fluid2d user script is written -- Param, Grid,
Fluid2d, model.var.get(...), model.set_psi_from_vorticity() -- against an `api`
namespace that carries those three classes, so the very same lines drive both
implementations.  The set-ups follow the reference experiments named in
BASELINE.json (sizes reduced so the CPU oracle finishes in seconds):

  freedecay : experiments/Twodim_turbulence/freedecay/freedecay.py
  vortex    : experiments/Vortex/vortex.py            (dipole2, closed, order 3)
  rb        : experiments/RayleighBenard/rayleigh_benard.py (+ coolroof forcing)
  karman    : experiments/VonKarman/karman_street.py  (island, noslip, sponge)
  qg        : experiments/QGbasic  (QG Helmholtz operator, LFAM3 stepper)
"""
import numpy as np


def _common(param, expname, datadir):
    param.expname = expname
    param.datadir = datadir
    param.plot_interactive = False
    param.generate_mp4 = False
    param.freq_his = 1e30
    param.freq_diag = 1e30
    param.exacthistime = False
    param.tend = 1e30
    param.npx = 1
    param.npy = 1
    param.tee_stdout = False   # (fluid2d_b200 extension: do not tee stdout into expdir)


def freedecay(api, datadir, n=64, order=5, tracer=True, timestepping='RK3_SSP', ny=None, npy=1, tile=False,
              diag_fluxes=False):
    """ny != n gives a rectangular domain with dx = dy; npy > 1 splits it in y-slabs (one
    rank per slab, like the reference run under mpirun with npx=1, npy=nranks); tile=True
    (needs ny = n*npy) repeats the n x n field in every slab instead of drawing a global one;
    tile=T (int) tiles a T x T field over the whole domain"""
    param = api.Param('default.xml')
    param.modelname = 'euler'
    _common(param, 'freedecay_%i' % n, datadir)
    param.nx = n
    param.ny = n if ny is None else ny
    param.Ly = param.Lx*param.ny/param.nx
    param.npy = npy
    param.geometry = 'perio'
    param.cfl = 1.2
    param.adaptable_dt = True
    param.dt = .05
    param.dtmax = 10.
    param.order = order
    param.timestepping = timestepping
    param.var_to_save = ['vorticity', 'psi', 'tracer'] if tracer else ['vorticity', 'psi']
    param.forcing = False
    param.noslip = False
    param.diffusion = False
    param.diag_fluxes = diag_fluxes
    if tracer:
        param.additional_tracer = ['tracer']
    grid = api.Grid(param)
    param.Kdiff = 5e-4*grid.dx
    f2d = api.Fluid2d(param, grid)
    model = f2d.model
    xr, yr = grid.xr, grid.yr
    vor = model.var.get('vorticity')

    np.random.seed(42)

    def wavenumbers(n, L):
        k = ((n//2+np.arange(n)) % n) - n//2
        return 2*np.pi*k/L

    # tile = True: the nx x nx field repeated in every slab; tile = T (an int): a T x T field
    # repeated periodically over the whole domain in both directions (large strong-scaling
    # grids: nobody has to hold a global 16384^2 complex spectrum on the host)
    both = not isinstance(tile, bool)
    fnx = int(tile) if both else param.nx
    fny = int(tile) if both else (param.nx if tile else param.ny)
    kkx, kky = np.meshgrid(wavenumbers(fnx, np.pi), wavenumbers(fny, np.pi))
    kk = np.sqrt(kkx**2 + kky**2)
    k0 = fnx*0.48
    dk = 1
    phase = np.random.normal(size=(fny, fnx))*2*np.pi
    hnoise = np.exp(-(kk-k0)**2/(2*dk))*np.exp(1j*phase)
    noise = np.zeros_like(vor)
    nh = grid.nh
    field = 1e3*np.real(np.fft.ifft2(hnoise))     # the global field, identical on every rank
    rows = param.ny//param.npy
    if both:
        jj = (grid.j0*rows+np.arange(rows)) % fny
        noise[nh:-nh, nh:-nh] = field[jj][:, np.arange(param.nx) % fnx]
    else:
        j0 = 0 if tile else grid.j0
        noise[nh:-nh, nh:-nh] = field[j0*rows:(j0+1)*rows, :]
    grid.fill_halo(noise)
    vor[:] = noise
    if tracer:
        trac = model.var.get('tracer')
        trac[:] = np.round(xr*6) % 2 + np.round(yr*6) % 2
    model.set_psi_from_vorticity()
    model.diagnostics(model.var, 0)
    model.set_psi_from_vorticity()
    return f2d


def vortex(api, datadir, n=64, order=3, msk_config='none', timestepping='RK3_SSP', diffusion=False, cfl=1.):
    param = api.Param('default.xml')
    param.modelname = 'euler'
    _common(param, 'vortex_%i_%s_%s' % (n, msk_config, timestepping), datadir)
    param.nx = n
    param.ny = n
    param.Ly = param.Lx
    param.geometry = 'closed'
    param.cfl = cfl
    param.adaptable_dt = True
    param.dt = 0.01
    param.dtmax = 100
    param.order = order
    param.timestepping = timestepping
    param.var_to_save = ['vorticity', 'psi', 'tracer']
    param.noslip = False
    param.diffusion = diffusion
    param.additional_tracer = ['tracer']
    grid = api.Grid(param)
    param.Kdiff = 5e-2*grid.dx
    xr, yr = grid.xr, grid.yr
    if msk_config == 'T-wall':
        i0, j0 = param.nx//2, param.ny//2
        di = int(0.25*param.Lx/grid.dx)
        grid.msk[:j0, i0] = 0
        grid.msk[j0, i0-di:i0+di] = 0
        grid.finalize_msk()
    f2d = api.Fluid2d(param, grid)
    model = f2d.model
    vor = model.var.get('vorticity')

    def gaussian(x0, y0, sigma):
        r2 = (xr-param.Lx*x0)**2+(yr-param.Ly*y0)**2
        return np.exp(-r2/(sigma**2))

    sigma = 0.05*param.Lx
    vor[:] = -gaussian(0.7, 0.42, sigma)
    vor[:] += gaussian(0.7, 0.58, sigma)
    vor[:] = vor*grid.msk
    model.set_psi_from_vorticity()
    state = model.var.get('tracer')
    state[:] = np.round(xr*6) % 2 + np.round(yr*6) % 2
    state *= grid.msk
    model.diagnostics(model.var, 0)
    enstrophy = model.diags['enstrophy']
    vor[:] = vor[:] / np.sqrt(enstrophy)
    model.set_psi_from_vorticity()
    return f2d


class CoolRoof(object):
    """user forcing in the style of experiments/RayleighBenard/forcing_rayleigh.py
    (type 'coolroof'): a buoyancy flux +Q on the bottom row, -Q on the top row; it
    scales the whole buoyancy tendency by `coef` like the reference script does."""

    def __init__(self, param, grid):
        Q = 1e-2
        nh = param.nh
        self.forc = grid.yr*0.
        self.forc[-nh-1, :] = -Q
        self.forc[nh, :] = +Q
        self.forc *= grid.msk
        self.forc *= (1./grid.dx)

    def add_forcing(self, x, t, dxdt, coef=1.):
        dxdt[4] += self.forc
        dxdt[4] *= coef


def rb(api, datadir, nx=64, diag_fluxes=False):
    param = api.Param('default.xml')
    param.modelname = 'boussinesq'
    _common(param, 'rb_%i' % nx, datadir)
    param.nx = nx
    param.ny = param.nx/2
    param.Lx = 2.
    param.Ly = 1.
    param.geometry = 'xchannel'
    param.cfl = 1.
    param.adaptable_dt = True
    param.dt = .1
    param.dtmax = .1
    param.order = 5
    param.aparab = 0.02
    param.var_to_save = ['vorticity', 'buoyancy', 'v', 'psi']
    param.gravity = 1.
    param.forcing = True
    param.forcing_module = 'embedded'
    param.diffusion = True
    param.noslip = True
    param.diag_fluxes = diag_fluxes
    grid = api.Grid(param)
    param.deltab = 600
    visco = .002*grid.dy
    param.Kdiff = {}
    param.Kdiff['vorticity'] = visco
    param.Kdiff['buoyancy'] = visco
    f2d = api.Fluid2d(param, grid)
    model = f2d.model
    model.forc = CoolRoof(param, grid)
    yr = grid.yr
    buoy = model.var.get('buoyancy')
    np.random.seed(1)
    noise = np.random.normal(size=np.shape(yr))*grid.msk
    noise -= grid.domain_integration(noise)*grid.msk/grid.area
    grid.fill_halo(noise)
    buoy += 1e-1*noise
    model.set_psi_from_vorticity()
    return f2d


class SaltFingerFluxes:
    """embedded forcing of experiments/doublediffusion/doublediffusion.py:70-112: constant
    temperature / salinity fluxes through the bottom and top rows of the channel"""

    def __init__(self, param, grid):
        self.nh = grid.nh
        dz = grid.dy
        self.FluxT = param.Kdiff['T']*param.dTdz/dz
        self.FluxS = param.Kdiff['S']*param.dSdz/dz

    def add_forcing(self, x, t, dxdt, coef=1.):
        dxdt[6][-self.nh-1, :] += self.FluxT
        dxdt[7][-self.nh-1, :] += self.FluxS
        dxdt[6][self.nh, :] -= self.FluxT
        dxdt[7][self.nh, :] -= self.FluxS


def dbldiff(api, datadir, nx=32, relaxation='default'):
    """experiments/doublediffusion/doublediffusion.py: the BoussinesqTS model (temperature +
    salinity, density diagnosed), tall channel, per-tracer diffusivities, flux forcing.  The
    script asks for relaxation='tridiagonal' (the line smoother of the multigrid)."""
    param = api.Param('default.xml')
    param.modelname = 'boussinesqTS'
    _common(param, 'dbldiff_%i' % nx, datadir)
    param.nx = nx
    param.ny = param.nx*2
    param.Lx = 1.
    param.Ly = param.Lx*2
    param.geometry = 'xchannel'
    param.cfl = 1.
    param.adaptable_dt = True
    param.dt = 0.01
    param.dtmax = 1e-2
    param.order = 5
    param.relaxation = relaxation
    param.var_to_save = ['vorticity', 'density', 'T', 'S', 'psi']
    param.gravity = 1.
    param.forcing = True
    param.forcing_module = 'embedded'
    param.diffusion = True
    param.noslip = False
    param.alphaT = 1.
    param.betaS = 1.
    grid = api.Grid(param)
    K0 = 2e-2*grid.dx
    param.Kdiff = {'vorticity': K0*7., 'T': K0, 'S': K0/50.}
    param.dTdz = 50.
    param.dSdz = 50.
    f2d = api.Fluid2d(param, grid)
    model = f2d.model
    model.forc = SaltFingerFluxes(param, grid)
    z = grid.yr0
    temp = model.var.get('T')
    salt = model.var.get('S')
    vor = model.var.get('vorticity')
    temp[:, :] = param.dTdz*z
    salt[:, :] = param.dSdz*z
    np.random.seed(42)
    noise = np.random.normal(size=np.shape(grid.yr), scale=1.)*grid.msk
    noise -= grid.domain_integration(noise)*grid.msk/grid.area
    grid.fill_halo(noise)
    vor[:, :] = 1e-2*noise
    model.set_density()
    model.set_psi_from_vorticity()
    return f2d


def karman(api, datadir, ny=32, ratio=2):
    param = api.Param('default.xml')
    param.modelname = 'euler'
    _common(param, 'karman_%i' % ny, datadir)
    param.ny = ny
    param.nx = param.ny*ratio
    param.Ly = 1.
    param.Lx = param.Ly*ratio
    param.geometry = 'xchannel'
    param.cfl = 1.2
    param.adaptable_dt = True
    param.dt = 1e-2
    param.dtmax = 1.
    param.order = 3
    param.timestepping = 'RK3_SSP'
    param.var_to_save = ['vorticity', 'psi', 'u']
    param.forcing = False
    param.noslip = True
    param.diffusion = True
    param.isisland = True
    param.spongelayer = True
    param.decay = False
    nh = param.nh
    grid = api.Grid(param)
    xr, yr = grid.xr, grid.yr
    psi0 = 0.2
    sigma = 0.16
    r = np.sqrt((xr-0.5*param.Ly)**2+(yr-0.5)**2)
    idx = np.where(r <= sigma)
    grid.msk[idx] = 0
    grid.msknoslip[idx] = 0
    grid.island.add(idx, 0.)
    msk = grid.msk.copy()*0
    msk[:nh, :] = 1
    grid.island.add(np.where(msk == 1), psi0*.5)
    msk = grid.msk.copy()*0
    msk[-nh:-1, :] = 1
    grid.island.add(np.where(msk == 1), -psi0*.5)
    grid.msknoslip[:nh, :] = 1
    grid.msknoslip[-nh:, :] = 1
    param.Kdiff = 5e-3*grid.dx
    f2d = api.Fluid2d(param, grid)
    model = f2d.model
    vor = model.var.get('vorticity')
    np.random.seed(1)
    noise = np.random.normal(size=np.shape(yr))*grid.msk
    grid.fill_halo(noise)
    noise -= grid.domain_integration(noise)*grid.msk/grid.area
    vor += 1e-1*noise*grid.msk
    vor *= grid.msk
    model.set_psi_from_vorticity()
    return f2d


def qg(api, datadir, n=64, timestepping='LFAM3', diagnosed=False):
    """diagnosed=True: with the bottom-torque and ageostrophic-velocity diagnostics
    (quasigeostrophic.py:96-153; experiments/QGbasic/vortex_betaplane_v2.py)"""
    param = api.Param('default.xml')
    param.modelname = 'quasigeostrophic'
    _common(param, 'qg_%i' % n, datadir)
    param.bottom_torque = diagnosed
    param.ageostrophic = diagnosed
    param.nx = n
    param.ny = n
    param.geometry = 'closed'
    param.cfl = 0.8
    param.adaptable_dt = True
    param.dt = 1.
    param.dtmax = 100.
    param.order = 5
    param.timestepping = timestepping
    param.var_to_save = ['pv', 'psi']
    param.beta = 1.
    param.Rd = 0.1
    param.forcing = False
    param.noslip = False
    param.diffusion = False
    grid = api.Grid(param)
    f2d = api.Fluid2d(param, grid)
    model = f2d.model
    xr, yr = grid.xr, grid.yr
    pv = model.var.get('pv')
    d2 = (xr-0.4)**2+(yr-0.55)**2
    pv[:] = 2.*np.exp(-d2/(0.08**2))*grid.msk
    model.add_backgroundpv()
    model.set_psi_from_pv()
    return f2d


def sqg(api, datadir, n=32):
    """experiments/SQG/vortex.py: surface QG, two co-rotating gaussian vortices, spectral
    inversion (the reference uses numpy.fft, the product cuFFT: agreement to FFT rounding)"""
    param = api.Param('default.xml')
    param.modelname = 'sqg'
    _common(param, 'sqg_%i' % n, datadir)
    param.nx = n
    param.ny = n
    param.geometry = 'perio'
    param.cfl = 0.8
    param.adaptable_dt = True
    param.dt = 1.
    param.dtmax = 100.
    param.order = 5
    param.ageostrophic = False
    param.var_to_save = ['pv', 'psi', 'u', 'v', 'vorticity']
    param.beta = 0.
    param.Rd = 10.
    param.forcing = False
    param.noslip = False
    param.diffusion = False
    grid = api.Grid(param)
    f2d = api.Fluid2d(param, grid)
    model = f2d.model
    xr, yr = grid.xr, grid.yr
    pv = model.var.get('pv')
    sigma = 0.1*param.Lx
    d = 3*sigma/param.Lx
    for y0 in (0.5-d/2, 0.5+d/2):
        r2 = (xr-param.Lx*0.5)**2+(yr-param.Ly*y0)**2
        pv[:] += -.5*np.exp(-r2/(sigma**2))
    model.set_psi_from_pv()
    return f2d


def symmetric_instab(api, datadir, ny=32):
    """experiments/SymmetricInstability/symmetric_instab.py: the thermal-wind model (vorticity,
    buoyancy, along-front velocity V, diagnosed Ertel PV), closed basin, order 3, with
    diag_fluxes (the script's 'symmetric' configuration)"""
    param = api.Param('default.xml')
    param.modelname = 'thermalwind'
    _common(param, 'si_%i' % ny, datadir)
    ratio = 2
    param.ny = ny
    param.nx = param.ny*ratio
    param.Ly = 1.
    param.Lx = param.Ly*ratio
    param.geometry = 'closed'
    param.cfl = 1.0
    param.adaptable_dt = True
    param.dt = .1
    param.dtmax = 2.
    param.order = 3
    param.timestepping = 'RK3_SSP'
    param.var_to_save = ['vorticity', 'psi', 'V', 'buoyancy', 'qE', 'tracer']
    param.diag_fluxes = True
    param.forcing = False
    param.noslip = False
    param.diffusion = False
    param.gravity = 1.
    param.f0 = 0.1
    param.additional_tracer = ['tracer']
    grid = api.Grid(param)
    param.Kdiff = 5e-1*grid.dx**2
    f2d = api.Fluid2d(param, grid)
    model = f2d.model
    yr = grid.yr
    vor = model.var.get('vorticity')
    buoy = model.var.get('buoyancy')
    V = model.var.get('V')
    trac = model.var.get('tracer')
    V0, N2, delta, sigma, alpha = .2, .15, .4, .4, 1.
    bback = N2*grid.yr0/param.Ly
    z = grid.yr0/delta
    x = (grid.xr0)/sigma
    V[:, :] = V0*(1+np.tanh(z))*(1-alpha*2*x*np.exp(-x**2))
    buoy[:, :] = (V0*param.f0) / np.cosh(z)**2 * (x+alpha*np.exp(-x**2)) + bback
    V *= grid.msk
    buoy *= grid.msk
    trac[:, :] = np.round(grid.xr*3) % 2
    vor[:, :] = 0.
    np.random.seed(1)
    noise = np.random.normal(size=np.shape(yr))*grid.msk
    noise -= grid.domain_integration(noise)*grid.msk/grid.area
    grid.fill_halo(noise)
    vor[:, :] += 1e-1*noise*grid.msk
    vor *= grid.msk
    model.set_psi_from_vorticity()
    model.compute_pv()
    return f2d


CASES = {
    'freedecay_64': lambda api, d: freedecay(api, d, 64),
    'freedecay_32_o3_notracer': lambda api, d: freedecay(api, d, 32, order=3, tracer=False),
    'vortex_64': lambda api, d: vortex(api, d, 64),
    'vortex_32_twall_o5': lambda api, d: vortex(api, d, 32, order=5, msk_config='T-wall'),
    'rb_64': lambda api, d: rb(api, d, 64),
    'karman_32': lambda api, d: karman(api, d, 32),
    'qg_32': lambda api, d: qg(api, d, 32),
    # diag_fluxes=True: the fixture also holds the reversible / irreversible flux stack of
    # core/fluxes.py at the initial state (flx0) and after the ten steps (flx10)
    'freedecay_32_flx': lambda api, d: freedecay(api, d, 32, diag_fluxes=True),
    'rb_32_flx': lambda api, d: rb(api, d, 32, diag_fluxes=True),
    # BoussinesqTS (temperature + salinity), damped-Jacobi and line (tridiagonal) relaxation
    'dbldiff_32': lambda api, d: dbldiff(api, d, 32),
    'dbldiff_32_tridiag': lambda api, d: dbldiff(api, d, 32, relaxation='tridiagonal'),
    'qg_32_diagnosed': lambda api, d: qg(api, d, 32, timestepping='RK3_SSP', diagnosed=True),
    'sqg_32': lambda api, d: sqg(api, d, 32),
    'si_32_flx': lambda api, d: symmetric_instab(api, d, 32),
}
# cases whose inversion is an FFT: another FFT library agrees to rounding, not bit for bit
SPECTRAL = {'sqg_32'}
# cases added after the last GPU session of round 1: their GPU parity test sits in
# tests/test_gpu_zz_late.py so that it runs after every test that has already been green on a
# B200 (the host layer is pinned on the CPU by tests/test_host_emulated.py)
LATE = {'dbldiff_32', 'dbldiff_32_tridiag', 'qg_32_diagnosed', 'sqg_32', 'si_32_flx'}
# every other stepper of core/timescheme.py:78-201, with diffusion on so that the
# `kstage == kforcing` branch of Euler.dynamics is exercised (light fixtures: states only)
SCHEMES = ['EF', 'LF', 'Heun', 'AB2', 'AB3', 'LFAM3', 'RK4_LS']   # ('RK3' is not in param.py's list)
LIGHT = set()
for _ts in SCHEMES:
    CASES['vortex_32_%s' % _ts] = (lambda api, d, ts=_ts: vortex(api, d, 32, timestepping=ts, diffusion=True, cfl=0.25))
    LIGHT.add('vortex_32_%s' % _ts)

NSTEPS = (1, 10)


def run_steps(f2d, nsteps=NSTEPS):
    """Advance with the same sequence Fluid2d.loop() performs per iteration
    (fluid2d.py:235-280: set_dt, step, t+=dt, diagnostics[, zero momentum]) and
    return {k: (state copy, t, dt, diags)} at the requested step counts."""
    out = {}
    model = f2d.model
    model.diagnostics(model.var, f2d.t)
    for kt in range(1, max(nsteps)+1):
        f2d.set_dt(f2d.kt)
        model.step(f2d.t, f2d.dt)
        f2d.t += f2d.dt
        f2d.kt += 1
        model.diagnostics(model.var, f2d.t)
        if f2d.enforce_momentum:
            f2d.enforce_zero_momentum()
            model.diagnostics(model.var, f2d.t)
        if kt in nsteps:
            out[kt] = (np.array(model.var.state, dtype=float, copy=True), float(f2d.t),
                       float(f2d.dt),
                       {k: float(np.asarray(v).ravel()[0]) for k, v in model.diags.items()})
    return out


def run_fluxes(f2d):
    """The flux diagnostic the way Fluid2d.loop() triggers it before a history write
    (fluid2d.py:204-210, 290-292): refresh maxspeed, set dt, Fluxes.diag_fluxes on the
    model state; returns a copy of the stack [rev_x_*, rev_y_*, ..., irr_x_*, ...]."""
    model = f2d.model
    model.diagnostics(model.var, f2d.t)
    f2d.set_dt(f2d.kt)
    f2d.flx.diag_fluxes(model.var.state, f2d.t, f2d.dt)
    return np.array(f2d.flx.flx, dtype=float, copy=True)
