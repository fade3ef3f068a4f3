"""Plumbing shared by the model classes (Euler, Boussinesq, QG): picking the attributes a
model needs from param / grid, publishing the layout of the state vector, and loading the
user-supplied classes (forcing, customized step) the reference API allows."""
from importlib import import_module


def adopt(obj, source, names):
    """obj.<name> = source.<name> for every name that source defines"""
    for name in names:
        if hasattr(source, name):
            setattr(obj, name, getattr(source, name))


def declare_state(param, grid, fields, tracers, whosetspsi, more_tracers=()):
    """Publish the state layout on param, where Var / Operators / Timescheme / Output read it:
    varname_list (order = field index on the device), tracer_list (what the advection
    kernel transports), whosetspsi (the field the inversion reads), sizevar."""
    param.varname_list = list(fields)
    param.tracer_list = list(tracers)
    for name in more_tracers:
        param.varname_list.append(name)
        param.tracer_list.append(name)
    param.whosetspsi = whosetspsi
    param.sizevar = [grid.nyl, grid.nxl]


def user_object(module_name, class_name, param, grid, what):
    """instance of class_name(param, grid) from a module on the python path
    (param.forcing_module -> Forcing, param.custom_module -> Step)"""
    try:
        module = import_module(module_name)
    except ImportError:
        raise ImportError('%s: cannot import module %r (is %s.py on the python path?)'
                          % (what, module_name, module_name))
    return getattr(module, class_name)(param, grid)


EMBEDDED_FORCING_NOTE = ('param.forcing is on with forcing_module = "embedded": the script must attach\n'
                         'its forcing itself, right after `model = f2d.model`:\n'
                         '    model.forc = Forcing(param, grid)')
