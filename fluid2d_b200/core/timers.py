"""Host-side stopwatch per phase of a step (reference interface: core/timers.py --
Timers(param), tic(name), toc(name), _print()).

The device queue is asynchronous: between tic and toc the host only ENQUEUES the kernels of
the phase, so the figures are host time unless sync=True, which brackets every phase with
a device synchronisation (for profiling only: it serialises host and device)."""
from time import perf_counter


class _Phase(object):
    __slots__ = ('started', 'total', 'calls')

    def __init__(self):
        self.started, self.total, self.calls = 0., 0., 0


class Timers(object):
    def __init__(self, param=None, sync=False):
        self.phases = {}
        self.sync = sync

    def _device_idle(self):
        if self.sync:
            import torch
            torch.cuda.synchronize()

    def tic(self, name):
        phase = self.phases.get(name)
        if phase is None:
            phase = self.phases[name] = _Phase()
        self._device_idle()
        phase.started = perf_counter()

    def toc(self, name):
        self._device_idle()
        phase = self.phases[name]
        phase.total += perf_counter()-phase.started
        phase.calls += 1

    # the reference exposes the accumulated figures as two dicts
    @property
    def elapse(self):
        return {name: p.total for name, p in self.phases.items()}

    @property
    def ncalls(self):
        return {name: p.calls for name, p in self.phases.items()}

    def _print(self):
        for name in sorted(self.phases):
            p = self.phases[name]
            if p.calls:
                print('%10s : %6.2f s / %6i calls / %6.2e' % (name, p.total, p.calls, p.total/p.calls))
