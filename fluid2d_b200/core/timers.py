"""tic/toc wall-clock timers per phase (reference: core/timers.py); on the device the
phases are asynchronous, so a toc only measures host time unless sync=True."""
from time import time as clock


class Timers(object):
    def __init__(self, param=None, sync=False):
        self.t0 = {}
        self.elapse = {}
        self.ncalls = {}
        self.sync = sync

    def _wait(self):
        if self.sync:
            import torch
            torch.cuda.synchronize()

    def tic(self, name):
        if name not in self.t0:
            self.elapse[name] = 0.
            self.ncalls[name] = 0.
        self._wait()
        self.t0[name] = clock()

    def toc(self, name):
        self._wait()
        self.elapse[name] += clock()-self.t0[name]
        self.ncalls[name] += 1

    def _print(self):
        for key in sorted(self.elapse):
            print('%10s : %6.2f s / %6i calls / %6.2e' %
                  (key, self.elapse[key], self.ncalls[key], self.elapse[key]/self.ncalls[key]))
