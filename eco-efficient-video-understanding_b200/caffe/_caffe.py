"""ctypes binding to libeco_b200.so -- the role boost.python's `_caffe.so` plays in the reference
(caffe_3d/python/caffe/_caffe.cpp:209-317).  Everything numeric happens behind the C ABI declared
in include/eco_b200.h; this file only marshals handles, shapes and fp32 host views."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("ECO_B200_LIB", os.path.join(_HERE, "..", "lib", "libeco_b200.so"))
_lib = None


class OpTime(C.Structure):
    _fields_ = [("name", C.c_char_p), ("kind", C.c_int), ("ms", C.c_float), ("flops", C.c_double),
                ("bytes", C.c_double)]


class ParamSlot(C.Structure):
    _fields_ = [("layer", C.c_int), ("blob", C.c_int), ("offset", C.c_size_t), ("count", C.c_size_t),
                ("lr_mult", C.c_float), ("decay_mult", C.c_float)]


GRAD_BUCKET_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_size_t, C.c_size_t)
GRAD_SYNC_FN = C.CFUNCTYPE(None, C.c_void_p)


def lib():
    """Load the CUDA library.  There is deliberately no fallback: without it nothing can run."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.abspath(_LIB_PATH)
    if not os.path.exists(path):
        raise RuntimeError("libeco_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (the product has no CPU / pure-Python path)" % path)
    L = C.CDLL(path)
    L.eco_last_error.restype = C.c_char_p
    L.eco_version.restype = C.c_char_p
    for fn in ("eco_net_name", "eco_net_layer_name", "eco_net_layer_type", "eco_net_blob_name"):
        getattr(L, fn).restype = C.c_char_p
    L.eco_net_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.eco_net_create_from_string.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.eco_net_create_from_string_until.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    L.eco_net_push_frames.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.eco_net_destroy.argtypes = [C.c_void_p]
    L.eco_net_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.eco_net_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.eco_net_copy_from.argtypes = [C.c_void_p, C.c_char_p]
    L.eco_net_save.argtypes = [C.c_void_p, C.c_char_p]
    L.eco_net_layer_num_params.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.eco_net_param_shape.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.eco_net_set_param.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_size_t]
    L.eco_net_get_param.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_size_t]
    L.eco_net_param_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_float)),
                                     C.POINTER(C.c_size_t)]
    for fn in ("eco_net_name", "eco_net_phase", "eco_net_num_layers", "eco_net_num_blobs", "eco_net_num_inputs",
               "eco_net_num_outputs"):
        getattr(L, fn).argtypes = [C.c_void_p]
    for fn in ("eco_net_layer_name", "eco_net_layer_type", "eco_net_layer_num_bottoms", "eco_net_layer_num_tops",
               "eco_net_blob_name", "eco_net_input_blob", "eco_net_output_blob"):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_int]
    for fn in ("eco_net_layer_bottom", "eco_net_layer_top"):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.eco_net_layer_index.argtypes = [C.c_void_p, C.c_char_p]
    L.eco_net_blob_index.argtypes = [C.c_void_p, C.c_char_p]
    L.eco_net_blob_shape.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.eco_blob_reshape.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int]
    L.eco_net_reshape.argtypes = [C.c_void_p]
    L.eco_net_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.eco_net_backward.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.eco_net_sync.argtypes = [C.c_void_p]
    L.eco_net_clear_param_diffs.argtypes = [C.c_void_p]
    L.eco_net_param_diff_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_size_t)]
    L.eco_net_param_arena.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.eco_net_grad_arena.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.eco_net_num_param_slots.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.eco_net_param_slot.argtypes = [C.c_void_p, C.c_int, C.POINTER(ParamSlot)]
    L.eco_net_params_updated_on_device.argtypes = [C.c_void_p]
    L.eco_net_cuda_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.eco_net_set_grad_bucket_hook.argtypes = [C.c_void_p, C.c_int, GRAD_BUCKET_FN, C.c_void_p]
    L.eco_net_num_grad_buckets.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.eco_net_grad_bucket.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.eco_solver_create.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.eco_solver_create_from_string.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
    L.eco_solver_destroy.argtypes = [C.c_void_p]
    L.eco_solver_net.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.eco_solver_iter.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.eco_solver_learning_rate.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.eco_solver_step.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    L.eco_solver_apply_update.argtypes = [C.c_void_p]
    L.eco_solver_set_grad_sync.argtypes = [C.c_void_p, GRAD_SYNC_FN, C.c_void_p, C.c_int]
    L.eco_solver_snapshot.argtypes = [C.c_void_p, C.c_char_p]
    L.eco_solver_restore.argtypes = [C.c_void_p, C.c_char_p]
    L.eco_blob_host_data.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_float)),
                                     C.POINTER(C.c_size_t)]
    L.eco_blob_host_diff.argtypes = L.eco_blob_host_data.argtypes
    L.eco_net_set_input_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.eco_blob_device_f32.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.eco_net_forward_pipelined.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
    L.eco_net_forward_pipelined_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_float), C.c_int, C.c_void_p,
                                               C.c_size_t, C.POINTER(C.c_int)]
    L.eco_net_wait.argtypes = [C.c_void_p, C.c_int]
    L.eco_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.eco_host_free.argtypes = [C.c_void_p]
    L.eco_net_last_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.eco_net_profile_forward.argtypes = [C.c_void_p, C.POINTER(OpTime), C.c_int, C.POINTER(C.c_int)]
    L.eco_net_profile_train.argtypes = [C.c_void_p, C.POINTER(OpTime), C.c_int, C.POINTER(C.c_int)]
    L.eco_net_describe_plan.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.eco_device_count.argtypes = [C.POINTER(C.c_int)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().eco_last_error().decode("utf-8", "replace"))


def _view(ptr, count, shape, owner):
    """Zero-copy fp32 numpy view of library-owned host memory; the view's base chain keeps the
    owning net alive (reference: _caffe.cpp:171-191, test_net.py:47-59)."""
    if count == 0:
        return np.zeros(shape, np.float32)
    buf = (C.c_float * count).from_address(C.addressof(ptr.contents))
    buf._owner = owner
    return np.frombuffer(buf, dtype=np.float32).reshape(shape)


class Blob(object):
    """Mirror of the Blob exposed by _caffe.cpp:252-263."""

    def __init__(self, net, index):
        self._net = net
        self._i = index

    @property
    def shape(self):
        dims = (C.c_int * 8)()
        nd = C.c_int(8)
        check(lib().eco_net_blob_shape(self._net._h, self._i, dims, C.byref(nd)))
        return tuple(dims[k] for k in range(nd.value))

    def _legacy(self, k):
        s = self.shape
        if len(s) > 4:
            raise RuntimeError("Cannot use legacy accessors on Blobs with > 4 axes.")  # blob.hpp:133-141
        s4 = tuple(s) + (1,) * (4 - len(s))
        return s4[k]

    num = property(lambda self: self._legacy(0))
    channels = property(lambda self: self._legacy(1))
    height = property(lambda self: self._legacy(2))
    width = property(lambda self: self._legacy(3))

    @property
    def count(self):
        return int(np.prod(self.shape)) if self.shape else 1

    def reshape(self, *dims):
        arr = (C.c_int * len(dims))(*[int(d) for d in dims])
        check(lib().eco_blob_reshape(self._net._h, self._i, arr, len(dims)))

    @property
    def data(self):
        p = C.POINTER(C.c_float)()
        n = C.c_size_t()
        check(lib().eco_blob_host_data(self._net._h, self._i, 1, C.byref(p), C.byref(n)))
        return _view(p, n.value, self.shape, self._net)

    @property
    def diff(self):
        """fp32 view of the gradient (synced from the device after Net.backward()).  To SEED a backward pass write the
        gradient with set_diff() or pass it to Net.backward(**{blob: array}) -- that is what marks it as newer on the host."""
        p = C.POINTER(C.c_float)()
        n = C.c_size_t()
        check(lib().eco_blob_host_diff(self._net._h, self._i, 0, C.byref(p), C.byref(n)))
        return _view(p, n.value, self.shape, self._net)

    def set_diff(self, arr):
        p = C.POINTER(C.c_float)()
        n = C.c_size_t()
        check(lib().eco_blob_host_diff(self._net._h, self._i, 1, C.byref(p), C.byref(n)))
        _view(p, n.value, self.shape, self._net)[...] = arr


class ParamBlob(object):
    """A layer parameter blob (`layer.blobs[i]`): data is a writable fp32 view."""

    def __init__(self, net, layer, idx):
        self._net, self._l, self._k = net, layer, idx

    @property
    def shape(self):
        dims = (C.c_int * 8)()
        nd = C.c_int(8)
        check(lib().eco_net_param_shape(self._net._h, self._l, self._k, dims, C.byref(nd)))
        return tuple(dims[k] for k in range(nd.value))

    @property
    def count(self):
        return int(np.prod(self.shape))

    @property
    def data(self):
        p = C.POINTER(C.c_float)()
        n = C.c_size_t()
        check(lib().eco_net_param_host(self._net._h, self._l, self._k, C.byref(p), C.byref(n)))
        return _view(p, n.value, self.shape, self._net)

    @property
    def diff(self):
        """parameter gradient (layer.blobs[i].diff), synced from the device gradient arena"""
        p = C.POINTER(C.c_float)()
        n = C.c_size_t()
        check(lib().eco_net_param_diff_host(self._net._h, self._l, self._k, C.byref(p), C.byref(n)))
        return _view(p, n.value, self.shape, self._net)

    num = property(lambda self: (tuple(self.shape) + (1, 1, 1, 1))[0])
    channels = property(lambda self: (tuple(self.shape) + (1, 1, 1, 1))[1])
    height = property(lambda self: (tuple(self.shape) + (1, 1, 1, 1))[2])
    width = property(lambda self: (tuple(self.shape) + (1, 1, 1, 1))[3])


class Layer(object):
    def __init__(self, net, index):
        self._net, self._i = net, index

    @property
    def type(self):
        return lib().eco_net_layer_type(self._net._h, self._i).decode()

    @property
    def blobs(self):
        n = C.c_int()
        check(lib().eco_net_layer_num_params(self._net._h, self._i, C.byref(n)))
        return [ParamBlob(self._net, self._i, k) for k in range(n.value)]
