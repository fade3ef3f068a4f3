"""Plotting: live figure of the run (reference: core/plotting.py).  Needs matplotlib,
which the GPU image does not ship; param.plot_interactive = False is the supported
mode there.  The class keeps the reference's call pattern: Plotting(param, grid, var,
diag), create_fig(t), update_fig(t, dt, kt), finalize()."""
import numpy as np


class Plotting(object):
    def __init__(self, param, grid, var, diag):
        try:
            import matplotlib
            matplotlib.use(getattr(param, 'mpl_backend', 'Agg'))
            import matplotlib.pyplot as plt
        except ImportError:
            raise ImportError('interactive plotting needs matplotlib; rerun with param.plot_interactive = False')
        self.plt = plt
        self.list_param = ['nh', 'plot_var', 'cax', 'colorscheme', 'cmap', 'expname', 'expdir', 'plot_psi']
        param.copy(self, self.list_param)
        self.grid, self.var, self.diag = grid, var, diag

    def create_fig(self, t):
        nh = self.nh
        self.fig, self.ax = self.plt.subplots()
        z = np.asarray(self.var.get(self.plot_var))[nh:-nh, nh:-nh]
        self.im = self.ax.imshow(z, origin='lower', cmap=self.cmap, interpolation='nearest')
        self.ax.set_title('%s / t=%.2f' % (self.plot_var, t))

    def update_fig(self, t, dt, kt):
        nh = self.nh
        z = np.asarray(self.var.get(self.plot_var))[nh:-nh, nh:-nh]
        self.im.set_array(z)
        if self.colorscheme == 'imposed':
            self.im.set_clim(self.cax)
        else:
            self.im.set_clim(z.min(), z.max())
        self.ax.set_title('%s / t=%.2f / kt=%i' % (self.plot_var, t, kt))
        self.fig.canvas.draw_idle()

    def finalize(self):
        self.fig.savefig('%s/%s.png' % (self.expdir, self.expname))
