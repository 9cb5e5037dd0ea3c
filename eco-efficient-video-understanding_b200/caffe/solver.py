"""Solver classes of pycaffe (caffe_3d/python/caffe/_caffe.cpp:296-317: Solver / SGDSolver / NesterovSolver with
`net`, `iter`, `step(n)`, `solve()`, `restore(path)`, and module-level get_solver) on top of the C ABI's eco_solver.

    solver = caffe.SGDSolver('solver.prototxt')           # `net:` is read next to the solver file
    solver.net.blobs['data'].data[...] = clips            # the caller feeds the train net's inputs
    solver.net.blobs['label'].data[...] = labels
    solver.step(1)

The reference's train net starts with a VideoData layer; that host-side pipeline is outside this hot path,
so the two tops it would produce are ordinary net inputs here."""
from __future__ import annotations

import ctypes as C
import os

from . import _caffe
from ._caffe import check, lib
from .pycaffe import Net


class _SolverNet(Net):
    """the solver's train net: a borrowed handle (the solver owns it)"""

    def __init__(self, handle, owner):
        self._h = handle
        self._owner = owner
        self._refresh_registry()

    def __del__(self):
        self._h = C.c_void_p()


class Solver(object):
    _TYPE = None

    def __init__(self, solver_file=None, solver_text=None, net_text=None, **net_options):
        L = lib()
        self._h = C.c_void_p()
        if solver_text is None:
            if not os.path.isfile(solver_file):
                raise RuntimeError("Could not open file " + str(solver_file))
            check(L.eco_solver_create(os.fsencode(solver_file), C.byref(self._h)))
        else:
            check(L.eco_solver_create_from_string(solver_text.encode(), net_text.encode() if net_text else None,
                                                  C.byref(self._h)))
        nh = C.c_void_p()
        check(L.eco_solver_net(self._h, C.byref(nh)))
        self.net = _SolverNet(nh, self)
        for k, v in net_options.items():
            self.net.set_option(k, v)
        self.test_nets = []
        self._sync_cb = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().eco_solver_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @property
    def iter(self):
        n = C.c_int()
        check(lib().eco_solver_iter(self._h, C.byref(n)))
        return n.value

    @property
    def learning_rate(self):
        r = C.c_float()
        check(lib().eco_solver_learning_rate(self._h, C.byref(r)))
        return r.value

    def step(self, n=1):
        loss = C.c_float()
        check(lib().eco_solver_step(self._h, int(n), C.byref(loss)))
        return loss.value

    def apply_update(self):
        check(lib().eco_solver_apply_update(self._h))

    def solve(self, resume_file=None, max_iter=None):
        if resume_file:
            self.restore(resume_file)
        if max_iter is None:
            raise RuntimeError("solve(): pass max_iter (the net's inputs are fed by the caller between steps; use step())")
        while self.iter < max_iter:
            self.step(1)

    def snapshot(self, prefix=None):
        check(lib().eco_solver_snapshot(self._h, prefix.encode() if prefix else None))

    def restore(self, state_file):
        if not os.path.isfile(state_file):
            raise RuntimeError("Could not open file " + str(state_file))
        check(lib().eco_solver_restore(self._h, os.fsencode(state_file)))

    def set_grad_sync(self, fn, world):
        """fn() is called between backward and update of every iteration (gradient exchange); world = replicas summed"""
        self._sync_cb = _caffe.GRAD_SYNC_FN(lambda user: fn())
        check(lib().eco_solver_set_grad_sync(self._h, self._sync_cb, None, int(world)))


class SGDSolver(Solver):
    pass


class NesterovSolver(Solver):
    pass


def get_solver(solver_file):
    """_caffe.cpp:312 GetSolverFromFile: the class follows the solver_type of the file"""
    return Solver(solver_file)
