"""Timescheme: catalog of time-stepping schemes around a user rhs(x, t, dxdt).

Same interface as the reference's core/timescheme.py (set(rhs, name), forward(x, t, dt),
kstage, kforcing, dtcoef, the persistent buffers x, dx0, dx1, dx2, xb whose psi slots
carry the first guess of the next truncated multigrid solve), with every whole-state
combination done by one kernel of libf2d_b200.so that reproduces the rounding sequence
of the numpy expression it replaces (f2d_ts_*).  x, dx* are DeviceState objects.
"""
from devarray import DeviceState
from runtime import rt


class Timescheme(object):
    def __init__(self, param, x):
        self.list_param = ['timestepping']
        param.copy(self, self.list_param)
        self.timeschemelist = {'EF': self.EulerForward, 'LF': self.LeapFrog, 'Heun': self.Heun,
                               'AB2': self.AB2, 'AB3': self.AB3, 'LFAM3': self.LFAM3,
                               'RK3_SSP': self.RK3_SSP, 'RK3': self.RK3, 'RK4_LS': self.RK4_LS}
        self.asselin_cst = 0.1
        self.ab2_epsilon = 0.1
        # weight of the terms evaluated only at the last stage of a multi-stage scheme
        coefondt = {'EF': 1., 'LF': 1., 'Heun': 2., 'AB2': 1., 'AB3': 1., 'LFAM3': 1.,
                    'RK3_SSP': 1.5, 'RK3': 1., 'RK4_LS': 1.}
        self.dtcoef = coefondt[self.timestepping]
        nvar, ny, nx = x.shape
        new = lambda: DeviceState(nvar, ny, nx)
        self.x = new()
        self.dx0 = new()
        self.dx1 = new()
        if self.timestepping in ['RK3_SSP', 'AB3', 'RK3']:
            self.dx2 = new()
        if self.timestepping in ['LF', 'LFAM3']:
            self.xb = new()
        self.first = True
        self.second = True
        self.n = x.size
        self.fieldsize = ny*nx
        # optional: the fields a model needs combined at the intermediate stages / at the
        # final update of RK3_SSP (None = the whole state, as the reference does).  A model
        # may only list fewer fields when skipping the others cannot change any value it
        # reads later (see Euler.__init__).
        self.fields_stage = None
        self.fields_final = None
        # optional: an object with `fuse` / `fused` attributes (the model's Operators) and the
        # fields whose FIRST-stage update x + dt*dx0 of RK3_SSP its advection kernel may write
        # itself -- tracers whose tendency is complete once rhs_adv has run at that stage
        self.adv_hook = None
        self.fused_fields = []
        # optional: the model's Operators and the indices (u, v): at the first two stages of
        # RK3_SSP the inversion that ends rhs() derives the velocity tendencies, and its
        # orthogradient kernel may write the stage velocities x[u] + c*(...) itself
        # (f2d_mg_set_uv_stage); only for models whose u, v tendencies are complete at that point
        self.uv_hook = None
        self.uv_fields = None
        self.kstage = 0
        self.kforcing = 0
        self.forward = self._unset

    def _unset(self, *args, **kwargs):
        raise RuntimeError('define a rhs and a timestepping with set() before calling forward()')

    def set(self, rhs, timestepping):
        self.rhs = rhs
        self.forward = self.timeschemelist[timestepping]
        self.kforcing = 0
        if self.timestepping in ['RK4_LS']:
            self.kforcing = 3
        elif self.timestepping in ['RK3_SSP']:
            self.kforcing = 2
        elif self.timestepping in ['Heun', 'LFAM3']:
            self.kforcing = 1

    # -- helpers: whole-state pointers ---------------------------------------
    @staticmethod
    def _r(s):
        return s.all_ptr(False)

    @staticmethod
    def _w(s):
        return s.all_ptr(True)

    def _runs(self, fields):
        """contiguous runs [(first field, number of fields)] of a sorted list of indices"""
        runs = []
        for k in sorted(set(fields)):
            if runs and runs[-1][0]+runs[-1][1] == k:
                runs[-1][1] += 1
            else:
                runs.append([k, 1])
        return runs

    @staticmethod
    def _rk(s, k0, cnt):
        p = s.rptr(k0)
        for k in range(k0+1, k0+cnt):
            s.rptr(k)
        return p

    @staticmethod
    def _wk(s, k0, cnt):
        p = s.wptr(k0)
        for k in range(k0+1, k0+cnt):
            s.wptr(k)
        return p

    def _copy(self, dst, src):
        r = rt()
        r.lib.copy(self._w(dst), self._r(src), self.n*8, r.stream)

    # -- schemes --------------------------------------------------------------
    def EulerForward(self, x, t, dt, **kwargs):
        r = rt()
        self.rhs(x, t, self.dx0)
        r.lib.ts_axpy(self._w(x), dt, self._r(self.dx0), self.n, r.stream)

    def AB2(self, x, t, dt):
        r = rt()
        self.rhs(x, t, self.dx0)
        if self.first:
            r.lib.ts_axpy(self._w(x), dt, self._r(self.dx0), self.n, r.stream)
            self.first = False
        else:
            r.lib.ts_ab2(self._w(x), (1.5+self.ab2_epsilon)*dt, self._r(self.dx0),
                         (0.5+self.ab2_epsilon)*dt, self._r(self.dx1), self.n, r.stream)
        self._copy(self.dx1, self.dx0)

    def AB3(self, x, t, dt, **kwargs):
        r = rt()
        self.rhs(x, t, self.dx0)
        if self.first:
            r.lib.ts_axpy(self._w(x), dt, self._r(self.dx0), self.n, r.stream)
            self.first = False
        elif self.second:
            r.lib.ts_ab2(self._w(x), 1.5*dt, self._r(self.dx0), 0.5*dt, self._r(self.dx1), self.n, r.stream)
            self.second = False
        else:
            r.lib.ts_ab3(self._w(x), 23*dt/12., self._r(self.dx0), 16*dt/12., self._r(self.dx1),
                         5*dt/12., self._r(self.dx2), self.n, r.stream)
        self._copy(self.dx2, self.dx1)
        self._copy(self.dx1, self.dx0)

    def LeapFrog(self, x, t, dt, **kwargs):
        r = rt()
        self._copy(self.x, x)
        self.rhs(x, t, self.dx0)
        if self.first:
            r.lib.ts_axpy(self._w(x), dt, self._r(self.dx0), self.n, r.stream)
            self.first = False
        else:
            r.lib.ts_set_xpay(self._w(x), self._r(self.xb), 2*dt, self._r(self.dx0), self.n, r.stream)
            r.lib.ts_asselin(self._w(self.x), self.asselin_cst, self._r(x), self._r(self.xb), self.n, r.stream)
        self._copy(self.xb, self.x)

    def LFAM3(self, x, t, dt, **kwargs):
        r = rt()
        self._copy(self.x, x)
        self.kstage = 0
        self.rhs(x, t, self.dx0)
        if self.first:
            r.lib.ts_axpy(self._w(x), dt, self._r(self.dx0), self.n, r.stream)
            self.first = False
        else:
            # leapfrog predictor to n+1, AM3 blend to n+1/2, corrector from there
            r.lib.ts_set_xpay(self._w(x), self._r(self.xb), 2*dt, self._r(self.dx0), self.n, r.stream)
            r.lib.ts_am3(self._w(x), self._r(self.x), self._r(self.xb), self.n, r.stream)
            self.kstage = 1
            self.rhs(x, t+dt*.5, self.dx0)
            r.lib.ts_set_xpay(self._w(x), self._r(self.x), dt, self._r(self.dx0), self.n, r.stream)
        self._copy(self.xb, self.x)

    def Heun(self, x, t, dt, **kwargs):
        r = rt()
        self.kstage = 0
        self.rhs(x, t, self.dx0)
        r.lib.ts_xpay(self._w(self.x), self._r(x), dt, self._r(self.dx0), self.n, r.stream)
        self.kstage = 1
        self.rhs(self.x, t+dt, self.dx1)
        r.lib.ts_xpay2(self._w(x), self._r(x), 0.5*dt, self._r(self.dx0), self._r(self.dx1), self.n, r.stream)

    def RK3(self, x, t, dt, **kwargs):
        r = rt()
        self.kstage = 0
        self.rhs(x, t, self.dx0)
        r.lib.ts_xpay(self._w(self.x), self._r(x), dt/3., self._r(self.dx0), self.n, r.stream)
        self.kstage = 1
        self.rhs(self.x, t+dt/3., self.dx1)
        r.lib.ts_xpay(self._w(self.x), self._r(x), 0.5*dt, self._r(self.dx1), self.n, r.stream)
        self.kstage = 2
        self.rhs(self.x, t+0.5*dt, self.dx2)
        r.lib.ts_axpy(self._w(x), dt, self._r(self.dx2), self.n, r.stream)

    def RK3_SSP(self, x, t, dt, **kwargs):
        r = rt()
        lib = r.lib
        nf = x.nvar
        stage = self._runs(self.fields_stage if self.fields_stage is not None else range(nf))
        final = self._runs(self.fields_final if self.fields_final is not None else range(nf))
        fs = self.fieldsize
        stage_fields = list(self.fields_stage if self.fields_stage is not None else range(nf))
        self.kstage = 0
        done = []
        uv = self.uv_hook if (self.uv_hook is not None and self.uv_fields is not None
                              and all(k in stage_fields for k in self.uv_fields)) else None
        if uv is not None:
            # (tendency buffer the inversion works on, base state, extra tendency, stage state, coefficient)
            uv.fuse_uv = (self.dx0, x, None, self.x, dt)
            uv.fused_uv = False
        try:
            if self.adv_hook is not None and self.fused_fields:
                self.adv_hook.fuse = (self.x, x, dt, list(self.fused_fields))
                try:
                    self.rhs(x, t, self.dx0)
                finally:
                    done = list(self.adv_hook.fused)
                    self.adv_hook.fuse = None
                    self.adv_hook.fused = []
            else:
                self.rhs(x, t, self.dx0)
        finally:
            if uv is not None:
                if uv.fused_uv:
                    done += list(self.uv_fields)
                uv.fuse_uv = None
        stage0 = self._runs([k for k in stage_fields if k not in done]) if done else stage
        for k0, c in stage0:
            lib.ts_xpay(self._wk(self.x, k0, c), self._rk(x, k0, c), dt, self._rk(self.dx0, k0, c), c*fs, r.stream)
        self.kstage = 1
        done1 = []
        if uv is not None:
            uv.fuse_uv = (self.dx1, x, self.dx0, self.x, 0.25*dt)
            uv.fused_uv = False
        try:
            self.rhs(self.x, t+dt, self.dx1)
        finally:
            if uv is not None:
                if uv.fused_uv:
                    done1 = list(self.uv_fields)
                uv.fuse_uv = None
        if done1:
            stage = self._runs([k for k in stage_fields if k not in done1])
        for k0, c in stage:
            lib.ts_xpay2(self._wk(self.x, k0, c), self._rk(x, k0, c), 0.25*dt, self._rk(self.dx0, k0, c),
                         self._rk(self.dx1, k0, c), c*fs, r.stream)
        self.kstage = 2
        self.rhs(self.x, t+0.5*dt, self.dx2)
        for k0, c in final:
            lib.ts_rk3ssp_final(self._wk(x, k0, c), dt/6., self._rk(self.dx0, k0, c), self._rk(self.dx1, k0, c),
                                self._rk(self.dx2, k0, c), c*fs, r.stream)

    def RK4_LS(self, x, t, dt, **kwargs):
        r = rt()
        for k, (c, tt) in enumerate([(0.25*dt, t), (dt/3., t+dt*0.25), (dt/2., t+dt/3.)]):
            self.kstage = k
            self.rhs(x if k == 0 else self.x, tt, self.dx0)
            r.lib.ts_xpay(self._w(self.x), self._r(x), c, self._r(self.dx0), self.n, r.stream)
        self.kstage = 3
        self.rhs(self.x, t+0.5*dt, self.dx0)
        r.lib.ts_axpy(self._w(x), dt, self._r(self.dx0), self.n, r.stream)
