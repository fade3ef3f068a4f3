"""Param: the attribute bag every Fluid2d script starts from.

Same surface as the reference's core/param.py:6-110 (Param(defaultfile), attribute
access, `avail` validation in checkall(), man()/manall(), listall(), printvalues(),
copy(obj, list_param) -> missing) and the same parameter names and default values as
core/defaults.json.  The defaults live in the table below rather than in a JSON file.
Unknown attributes may be added freely by scripts (param.gravity, param.tend, ...).
"""
import sys

# group -> name -> (default, allowed values or None, one-line help)
TABLE = {
    'general': {
        'modelname': ('advection', ['advection', 'euler', 'boussinesq', 'boussinesqTS',
                                    'quasigeostrophic', 'sqg', 'thermalwind'],
                      'set of equations that is integrated'),
        'expname': ('myexp', None, 'experiment name, used for the output directory and files'),
    },
    'numerics': {
        'timestepping': ('RK3_SSP', ['EF', 'LF', 'Heun', 'RK3_SSP', 'AB2', 'AB3', 'LFAM3', 'RK4_LS'],
                         'time scheme'),
        'order': (5, None, 'order of the flux interpolation: odd = upwind (1,3,5), even = centred (2,4,6)'),
        'aparab': (0.05, None, 'width of the parabolic flux splitting, as a fraction of max|u|'),
        'flux_splitting_method': ('parabolic', ['minmax', 'parabolic'], 'how |u| is regularised near u=0'),
        'relaxation': ('default', ['default', 'tridiagonal'], 'multigrid smoother'),
        'nh': (3, None, 'halo width (3 is compulsory)'),
    },
    'time': {
        'adaptable_dt': (True, None, 'dt follows the cfl criterion'),
        'dt': (0.1, None, 'time step (initial / fixed)'),
        'cfl': (0.5, None, 'target cfl number when adaptable_dt'),
        'dtmax': (5.0, None, 'upper bound of dt'),
        'rescaledtime': ('none', ['none', 'enstrophy'], 'rescale the model time'),
        'ninterrestart': (1, None, 'number of restarts a run is split in'),
    },
    'domain and resolution': {
        'nx': (128, None, 'number of cells in x (power of two)'),
        'ny': (128, None, 'number of cells in y (power of two)'),
        'Lx': (1.0, None, 'domain length in x'),
        'Ly': (1.0, None, 'domain length in y'),
        'geometry': ('disc', ['disc', 'perio', 'closed', 'ychannel', 'xchannel'], 'domain shape'),
        'isisland': (False, None, 'the domain has islands (multiply connected)'),
        'mpi': (0, None, 'legacy flag'),
        'myrank': (0, None, 'rank of this process'),
        'npx': (1, None, 'number of subdomains in x (power of two)'),
        'npy': (1, None, 'number of subdomains in y (power of two)'),
    },
    'plotting options': {
        'plot_interactive': (True, None, 'live figure during the run'),
        'imshow_interpolation': ('nearest', ['nearest', 'bilinear'], 'imshow interpolation'),
        'plot_psi': (False, None, 'overlay streamfunction contours'),
        'plot_ua': (False, None, 'overlay ageostrophic velocity'),
        'plot_pvback': (False, None, 'overlay background pv'),
        'freq_plot': (10, None, 'refresh the figure every freq_plot iterations'),
        'generate_mp4': (False, None, 'pipe the figure to ffmpeg'),
        'colorscheme': ('minmax', ['minmax', 'symmetric', 'imposed'], 'colour axis policy'),
        'cmap': ('RdBu_r', None, 'colormap'),
        'plotting_module': ('plotting', None, 'module that provides Plotting'),
    },
    'output': {
        'datadir': ('~/data/fluid2d', None, 'root of the output directories'),
        'expdir': ('none', None, 'set by Fluid2d: datadir/expname'),
        'var_to_save': ('vorticity', None, 'variables stored in the history file'),
        'list_diag': ('all', None, 'integral diagnostics stored in the diag file'),
        'nprint': (20, None, 'print the clock every nprint iterations'),
        'freq_his': (1.0, None, 'model time between two history snapshots'),
        'diag_fluxes': (False, None, 'diagnose reversible / irreversible fluxes'),
        'exacthistime': (True, None, 'trim dt to land on the history times'),
        'freq_diag': (1.0, None, 'model time between two diagnostics records'),
    },
    'physics': {
        'hydroepsilon': (1.0, None, 'aspect-ratio factor of the elliptic operator'),
        'diffusion': (False, None, 'add a Laplacian diffusion on the tracers'),
        'customized': (False, None, 'call a user Step after each time step'),
        'Kdiff': (0.0, None, 'diffusion coefficient (scalar or dict per tracer)'),
        'noslip': (False, None, 'no-slip boundary condition'),
        'ageostrophic': (False, None, 'QG: diagnose the ageostrophic velocity'),
        'bottom_torque': (False, None, 'QG: diagnose the bottom torque'),
        'forcing': (False, None, 'add a user forcing'),
        'forcing_module': ('embedded', None, "module providing Forcing, or 'embedded'"),
        'decay': (True, None, 'the kinetic energy is expected to decay'),
        'enforce_momentum': (False, None, 'remove the net momentum (closed/disc Euler)'),
        'spongelayer': (False, None, 'sponge at the eastern boundary (Euler)'),
    },
}


class Param(object):
    def __init__(self, defaultfile=None):
        # `defaultfile` is accepted and ignored, as in the reference (param.py:20-28)
        self.avail = {}
        self.doc = {}
        for group in TABLE.values():
            for name, (default, avail, doc) in group.items():
                setattr(self, name, default)
                if avail is not None:
                    self.avail[name] = avail
                self.doc[name] = doc
        args = sys.argv[1:]
        if '-h' in args:
            self.manall()
            sys.exit()
        self.print_param = '-v' in args

    def man(self, name):
        txt = self.doc.get(name, 'no manual for this parameter')
        if name in self.avail:
            txt += ' / available values = [' + ', '.join(str(a) for a in self.avail[name]) + ']'
        print('  - "\033[0;32;40m%s\033[0m" : %s\n' % (name, txt))

    def manall(self):
        for p in self.listall():
            self.man(p)

    def checkall(self):
        for p, avail in self.avail.items():
            if getattr(self, p) not in avail:
                raise ValueError('parameter "%s" should in %s' % (p, str(avail)))

    def listall(self):
        return [d for d in self.__dict__ if d not in ('avail', 'doc')]

    def printvalues(self):
        for d in self.listall():
            print('%20s :' % d, getattr(self, d))

    def copy(self, obj, list_param):
        """copy the attributes named in list_param onto obj; return the missing names"""
        missing = []
        for k in list_param:
            if hasattr(self, k):
                setattr(obj, k, getattr(self, k))
            else:
                missing.append(k)
        return missing
