"""Fluid2d: builds the model named by param.modelname and runs the time loop.

Same interface as the reference's core/fluid2d.py -- Fluid2d(param, grid), .model,
.loop(), .set_dt(kt), .enforce_zero_momentum(), .t, .kt, .dt, the stdout tee into
expdir/output.txt, the copy of the launch script, the blow-up stop, the end-of-run
performance summary (with the reference's "rescaled time") -- on top of the device
models.  During the loop the state never leaves HBM: per iteration the host reads back
eight diagnostics scalars (maxspeed sets the next dt) and the residual norms of the
end-of-step solve; fields cross only when Output stores a snapshot or a user hook asks.
"""
from __future__ import print_function
import os
import signal
import sys
from importlib import import_module
from subprocess import call
from time import time as clock

import numpy as np

from output import Output

# modelname -> (module, class); the models that need an FFT or the line relaxation are not
# on the device path
MODELS = {'euler': ('euler', 'Euler'), 'advection': ('advection', 'Advection'),
          'boussinesq': ('boussinesq', 'Boussinesq'), 'quasigeostrophic': ('quasigeostrophic', 'QG'),
          'boussinesqTS': ('boussinesqTS', 'BoussinesqTS'), 'sqg': ('sqg', 'SQG'),
          'thermalwind': ('thermalwind', 'Thermalwind')}
NOT_ON_DEVICE = ()
PV_MODELS = ('quasigeostrophic', 'sqg')

FROM_PARAM = ('modelname', 'tend', 'dt', 'adaptable_dt', 'cfl', 'dtmax', 'exacthistime', 'rescaledtime',
              'nprint', 'print_param', 'myrank', 'nbproc', 'npx', 'npy', 'nx', 'ny', 'geometry', 'noslip',
              'forcing', 'decay', 'enforce_momentum', 'isisland', 'diag_fluxes', 'plot_interactive',
              'plotting_module', 'freq_plot', 'freq_save')
FROM_GRID = ('dx', 'dy', 'nh', 'msk', 'xr0', 'yr0', 'x2', 'y2')
RULE = '-'*50


class Fluid2d(object):
    def __init__(self, param, grid):
        param.checkall()
        self._prepare_expdir(param)
        param.copy(self, FROM_PARAM)
        self.dt0 = self.dt
        grid.finalize_msk()
        grid.copy(self, FROM_GRID)
        self.grid = grid

        name = param.modelname
        if name in NOT_ON_DEVICE:
            raise NotImplementedError('model %s is outside the device hot path (see DESIGN.md)' % name)
        # the zero-momentum correction only makes sense for Euler in a closed basin
        if name != 'euler' or self.geometry not in ('closed', 'disc'):
            self.enforce_momentum = False
        if name in MODELS:
            module, cls = MODELS[name]
            self.model = getattr(import_module(module), cls)(param, grid)
        self.enstrophyname = 'pv2' if name in PV_MODELS else 'enstrophy'

        if self.isisland:
            grid.island.finalize(self.model.ope.mskp)
            self.model.ope.rhsp = grid.island.rhsp
            self.model.ope.psi = grid.island.psi
        flxlist = None
        if self.diag_fluxes:
            from fluxes import Fluxes
            self.flx = Fluxes(param, grid, self.model.ope)
            flxlist = self.flx.fullflx_list
        if self.plot_interactive:
            plotting = import_module(self.plotting_module)
            self.plotting = plotting.Plotting(param, grid, self.model.var, self.model.diags)
        self.tracer_list = param.tracer_list
        self.t = 0.
        self.kt = 0
        self.output = Output(param, grid, self.model.diags, flxlist=flxlist)
        self.print_config(param, start=True)

    def _prepare_expdir(self, param):
        """experiment directory, stdout tee, copy of the launch script (rank 0)"""
        launchscript = sys.argv[0]
        param.datadir = param.datadir.replace('~', os.getenv("HOME", '.'))
        param.expdir = '%s/%s' % (param.datadir, param.expname)
        if param.myrank != 0:
            return
        if not os.path.isdir(param.expdir):
            os.makedirs(param.expdir)
        outfile = '%s/output.txt' % param.expdir
        if os.path.exists(outfile):
            print('Warning: this experiment has already been ran, output.txt already exists')
            print('dummy.txt will be used instead')
            outfile = '%s/dummy.txt' % param.expdir
        if getattr(param, 'tee_stdout', True):
            sys.stdout = Logger(outfile)
        savedscript = '%s/%s.py' % (param.expdir, param.expname)
        if os.path.exists(savedscript):
            print('Warning: the python script already exists in %s' % param.expdir)
        elif os.path.isfile(launchscript):
            self.savedscript = savedscript
            call(['cp', launchscript, savedscript])

    @property
    def state(self):
        return self.model.var.state

    def print_config(self, param, start=True):
        if self.myrank != 0:
            return
        if start:
            print(RULE)
            print(' Fluid2d summary (B200 device build):')
            print(RULE)
            print('  - model equations: %s' % self.modelname)
            print('  - grid size: %i x %i' % (self.nx, self.ny))
            print('  - integration time: %.2f' % self.tend)
            print('  - advection schemes applied to:')
            for trac in self.tracer_list:
                print('    - %s' % trac)
            if self.print_param:
                param.printvalues()
            return
        files = [self.output.hisfile, self.output.diagfile]
        if self.diag_fluxes:
            files.append(self.output.flxfile)
        if hasattr(self, 'savedscript'):
            files.append(self.savedscript)
        print(' Output files:')
        print(RULE)
        for f in files:
            print('  - %s' % f)
        print(RULE)

    # ------------------------------------------------------------------ time loop
    def loop(self, joinhis=True, keepplotalive=False, nsteps=None):
        """time loop (fluid2d.py:188-349); nsteps (extension) stops after that many iterations"""
        if self.myrank == 0:
            print(RULE)
            print(' Starting the time loop')
            print(RULE)
        model = self.model
        model.diagnostics(model.var, self.t)
        model.diags['dkedt'] = 0.
        model.diags['dvdt'] = 0.
        data = {'his': model.var, 'diag': model.diags}
        if self.diag_fluxes:
            data['flx'] = self.flx
            # dt must be set first (adaptable_dt), fluid2d.py:204-210
            self.set_dt(self.kt)
            self.flx.diag_fluxes(model.var.dstate, self.t, self.dt)
        self.output.do(data, self.t, self.kt)
        if self.plot_interactive and not hasattr(self.plotting, 'fig'):
            self.plotting.create_fig(self.t)
        self._catch_ctrl_c()
        self.stop = False
        kt0 = self.kt
        t0 = clock()
        self._reduce = 0
        while self.t < self.tend and not self.stop:
            self.set_dt(self.kt)
            if self.exacthistime and self.adaptable_dt:
                self._land_on_history_time()
            model.step(self.t, self.dt)
            if self.rescaledtime == 'enstrophy':
                self.t += self.dt * np.sqrt(model.diags['enstrophy'])
            else:
                self.t += self.dt
            self.kt += 1
            self._diagnose_step(model)
            if self.diag_fluxes and (self.t >= self.output.tnexthis):
                # costly (two extra time steps): only before it is written
                self.flx.diag_fluxes(model.var.dstate, self.t, self.dt)
            self.output.do(data, self.t, self.kt)
            self._report_step(model)
            if nsteps is not None and self.kt-kt0 >= nsteps:
                break
        if self.myrank == 0:
            print('\ndone')
        if self.plot_interactive and not keepplotalive:
            self.plotting.finalize()
        if self.myrank == 0:
            self._print_performance(clock, t0, max(self.kt-kt0, 1))
            if joinhis:
                self.output.dump_diag()
                self.output.join()
        self.print_config(None, start=False)

    def _catch_ctrl_c(self):
        def handler(sig, frame):
            if self.myrank == 0:
                print('\n hit ctrl-C, stopping', end='')
            self.stop = True
        try:
            signal.signal(signal.SIGINT, handler)
        except ValueError:
            pass   # not in the main thread

    def _land_on_history_time(self):
        """shorten dt over the 8 steps before a history time so that a snapshot falls exactly
        on it (fluid2d.py:243-254)"""
        tnext = self.output.tnexthis
        if (self.t+8*self.dt > tnext) and (self._reduce == 0):
            self._reduce = 8
        if self._reduce > 0:
            self.dt = (tnext-self.t)/(self._reduce*0.95)
            self._reduce -= 1
        if self.t+self.dt > tnext:
            self._reduce = 0
            self.dt = tnext-self.t

    def _diagnose_step(self, model):
        """integral diagnostics of the new state and their rates of change"""
        ke_old = model.diags['ke']
        ens_old = model.diags[self.enstrophyname]
        model.diagnostics(model.var, self.t)
        if self.enforce_momentum:
            self.enforce_zero_momentum()
            model.diagnostics(model.var, self.t)
        ke = model.diags['ke']
        ens = model.diags[self.enstrophyname]
        model.diags['dkedt'] = (ke-ke_old)/self.dt
        model.diags['dvdt'] = (ens-ens_old)/self.dt
        if (ke > ke_old) and (self.myrank == 0) and self.decay and (self.modelname == 'euler'):
            print('\rkt=%-4i \033[0;32;40mWARNING dlog(ke)\033[0m = %.2g' %
                  (self.kt, float(np.ravel((ke-ke_old)/ke)[0])), end='')

    def _report_step(self, model):
        flag = '*' if self.dt == self.dtmax else ''
        if (self.myrank == 0) and (self.kt % self.nprint == 0) or (self.t >= self.tend):
            print('\rkt=%-4i / t=%-7.3f %s / dt=%-7.3f ' % (self.kt, self.t, flag, self.dt), end='')
        if self.plot_interactive and (self.kt % self.freq_plot == 0):
            self.plotting.update_fig(self.t, self.dt, self.kt)
        if model.diags['maxspeed'] > 1e3:
            self.stop = True
            if self.myrank == 0:
                print()
                print('max|u| > 1000, blow-up detected, stopping')

    def _print_performance(self, clock, t0, nkt):
        import torch
        torch.cuda.synchronize()
        wall = clock()-t0
        model = self.model
        if hasattr(model, 'timers'):
            print(RULE)
            print(' A few model performances metrics')
            print(RULE)
            model.timers._print()
        print()
        print('  - Wall  time      : %f s' % wall)
        print('  - Nb of iterations: %i' % nkt)
        print('  - Time per ite    : %5.3f s' % (wall/nkt))
        print('  - Rescaled time   : %5.3e s (per ite, per dof)' % (wall*self.npx*self.npy/(nkt*self.nx*self.ny)))
        print('  - Cell updates/s  : %5.3e' % (nkt*self.nx*self.ny/wall))
        print(RULE)

    # ------------------------------------------------------------------ helpers of the loop
    def enforce_zero_momentum(self):
        if self.enforce_momentum:
            model = self.model
            r = model.rt
            s = model.var.dstate
            px = float(np.ravel(model.diags['px']/self.x2)[0])
            py = float(np.ravel(model.diags['py']/self.y2)[0])
            r.lib.sub_lin2(s.wptr(model.var.index('vorticity')), px, r.ptr(model.ope.d_xr0), py,
                           r.ptr(model.ope.d_yr0), s.ny*s.nx, r.stream)
            model.ope.invert_vorticity(s, flag='fast')

    def set_dt(self, kt):
        maxspeed = self.model.diags['maxspeed']
        if ((self.adaptable_dt) & (maxspeed != 0)):
            dt = self.cfl * min(self.dx, self.dy) / maxspeed
            self.dt = dt
            if self.dt > self.dtmax:
                self.dt = self.dtmax
        else:
            self.dt = self.dt0
        # the advection scheme needs max|u| for the parabolic flux splitting
        self.model.ope.cst[3] = maxspeed


class Logger(object):
    """tee of stdout into expdir/output.txt"""

    def __init__(self, logfile):
        self.terminal = sys.stdout if not isinstance(sys.stdout, Logger) else sys.stdout.terminal
        self.log = open(logfile, "w")

    def write(self, message):
        self.terminal.write(message)
        self.log.write(message)

    def flush(self):
        self.terminal.flush()
