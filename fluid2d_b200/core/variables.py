"""Var: the packed model state (reference: core/variables.py).  state[nvar, ny, nx]
lives on the device (devarray.DeviceState); get(name) and .state hand out host views
that report their writes, so user scripts keep the reference's idiom

    vor = model.var.get('vorticity'); vor[:] = ...; model.set_psi_from_vorticity()
"""
from devarray import DeviceState


class Var(object):
    def __init__(self, param):
        self.list_param = ['sizevar', 'varname_list']
        param.copy(self, self.list_param)
        self.nvar = len(self.varname_list)
        if type(self.sizevar) != list:
            raise TypeError('sizevar has to be a list')
        self.sizestate = [self.nvar]+self.sizevar
        if getattr(param, 'npy', 1) > 1:
            from runtime import rt
            rt().ensure_comm(param.npy, self.sizevar[0]*self.sizevar[1]*8)
        self.dstate = DeviceState(self.nvar, self.sizevar[0], self.sizevar[1])

    @property
    def state(self):
        """writable host view of the whole state (device refreshed first)"""
        return self.dstate.host_view(None)

    def get(self, name):
        """writable host view of variable `name`"""
        return self.dstate.host_view(self.varname_list.index(name))

    def index(self, name):
        return self.varname_list.index(name)
