"""Pythonic Net on top of the C ABI: the surface of the reference's pycaffe
(caffe_3d/python/caffe/pycaffe.py:21-281 over _caffe.cpp:209-317), so scripts such as
scripts/online_recognition/online_recognition.py run unchanged with this package on sys.path:

    net = caffe.Net(prototxt, caffemodel, caffe.TEST)
    net.blobs['data'].data[...] = frames
    out = net.forward()                      # {'fc8': ndarray}
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict

import numpy as np

from . import _caffe
from ._caffe import Blob, Layer, check, lib

TRAIN = 0  # caffe.proto Phase
TEST = 1


class Net(object):
    """caffe.Net(prototxt, phase) / caffe.Net(prototxt, caffemodel, phase)  (_caffe.cpp:89-108)."""

    def __init__(self, *args, **kwargs):
        if len(args) == 2:
            model_file, phase = args
            weights = None
        elif len(args) == 3:
            model_file, weights, phase = args
        else:
            raise TypeError("Net(prototxt, phase) or Net(prototxt, caffemodel, phase)")
        self._h = C.c_void_p()
        text = kwargs.get("prototxt_text")
        if text is None:
            # same file check as the reference, which raises RuntimeError (_caffe.cpp:57-64)
            if not os.path.isfile(model_file):
                raise RuntimeError("Could not open file " + str(model_file))
            check(lib().eco_net_create(os.fsencode(model_file), int(phase), C.byref(self._h)))
        elif kwargs.get("until_blob"):
            check(lib().eco_net_create_from_string_until(text.encode(), int(phase), kwargs["until_blob"].encode(), C.byref(self._h)))
        else:
            check(lib().eco_net_create_from_string(text.encode(), int(phase), C.byref(self._h)))
        for k, v in kwargs.get("options", {}).items():
            check(lib().eco_net_set_option(self._h, k.encode(), int(v)))
        if weights is not None:
            if not os.path.isfile(weights):
                raise RuntimeError("Could not open file " + str(weights))
            self.copy_from(weights)
        self._refresh_registry()

    @classmethod
    def from_string(cls, text, phase, until_blob=None, **options):
        return cls("<string>", phase, prototxt_text=text, options=options, until_blob=until_blob)

    def push_frames(self, dst_blob, src_net, src_blob):
        """append the frames of src_net.blobs[src_blob] to the sliding window held in self.blobs[dst_blob] (device to device)"""
        check(lib().eco_net_push_frames(self._h, self._blob_names.index(dst_blob), src_net._h, src_net._blob_names.index(src_blob)))

    def _refresh_registry(self):
        L = lib()
        self._layer_names = [L.eco_net_layer_name(self._h, i).decode() for i in range(L.eco_net_num_layers(self._h))]
        self._blob_names = [L.eco_net_blob_name(self._h, i).decode() for i in range(L.eco_net_num_blobs(self._h))]
        self._blobs = [Blob(self, i) for i in range(len(self._blob_names))]
        self.layers = [Layer(self, i) for i in range(len(self._layer_names))]
        self._inputs = [L.eco_net_input_blob(self._h, i) for i in range(L.eco_net_num_inputs(self._h))]
        self._outputs = [L.eco_net_output_blob(self._h, i) for i in range(L.eco_net_num_outputs(self._h))]

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().eco_net_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- registry -------------------------------------------------------------------
    @property
    def name(self):
        return lib().eco_net_name(self._h).decode()

    @property
    def blobs(self):
        return OrderedDict(zip(self._blob_names, self._blobs))

    @property
    def params(self):
        out = OrderedDict()
        for name, lr in zip(self._layer_names, self.layers):
            b = lr.blobs
            if len(b) > 0:
                out[name] = b
        return out

    @property
    def inputs(self):
        return [self._blob_names[i] for i in self._inputs]

    @property
    def outputs(self):
        return [self._blob_names[i] for i in self._outputs]

    # ---- execution ------------------------------------------------------------------
    def _forward(self, start, end):
        loss = C.c_float(0)
        check(lib().eco_net_forward(self._h, int(start), int(end), C.byref(loss)))
        return loss.value

    def _backward(self, start, end):
        check(lib().eco_net_backward(self._h, int(start), int(end)))

    def forward(self, blobs=None, start=None, end=None, **kwargs):
        blobs = list(blobs or [])
        start_ind = self._layer_names.index(start) if start is not None else 0
        if end is not None:
            end_ind = self._layer_names.index(end)
            outputs = set([end] + blobs)
        else:
            end_ind = len(self.layers) - 1
            outputs = set(self.outputs + blobs)
        if kwargs:
            if set(kwargs.keys()) != set(self.inputs):
                raise Exception("Input blob arguments do not match net inputs.")
            allb = self.blobs
            for in_, arr in kwargs.items():
                if arr.shape[0] != allb[in_].shape[0]:
                    raise Exception("Input is not batch sized")
                allb[in_].data[...] = arr
        self._forward(start_ind, end_ind)
        allb = self.blobs
        return {out: allb[out].data for out in outputs}

    def backward(self, diffs=None, start=None, end=None, **kwargs):
        """pycaffe.py:117-160: backward pass; kwargs = {top blob: diff} seeds, returns {blob: diff} for the net inputs
        that carry a gradient plus the blobs named in `diffs`.  TRAIN-phase nets only."""
        diffs = list(diffs or [])
        start_ind = self._layer_names.index(start) if start is not None else len(self.layers) - 1
        if end is not None:
            end_ind = self._layer_names.index(end)
            outputs = set([end] + diffs)
        else:
            end_ind = 0
            outputs = set(self.inputs + diffs)
        allb = self.blobs
        if kwargs:
            for top, arr in kwargs.items():
                if top not in allb:
                    raise Exception("Top diff arguments do not match net outputs / blobs.")
                if arr.shape[0] != allb[top].shape[0]:
                    raise Exception("Diff is not batch sized")
                allb[top].set_diff(arr)
        self._backward(start_ind, end_ind)
        out = {}
        for name in outputs:
            if name in allb:
                try:
                    out[name] = allb[name].diff
                except RuntimeError:
                    pass
        return out

    def clear_param_diffs(self):
        check(lib().eco_net_clear_param_diffs(self._h))

    def forward_backward_all(self, blobs=None, diffs=None, **kwargs):
        """pycaffe.py:186-240, batch-wise forward + backward; kwargs hold inputs (and optionally top diffs) by name."""
        all_outs = {out: [] for out in set(self.outputs + list(blobs or []))}
        all_diffs = {d: [] for d in set(self.inputs + list(diffs or []))}
        fkw = {k: v for k, v in kwargs.items() if k in self.inputs}
        bkw = {k: v for k, v in kwargs.items() if k in self.outputs}
        n = len(next(iter(fkw.values())))
        fb = list(self._batch(fkw))
        bb = list(self._batch(bkw)) if bkw else [{}] * len(fb)
        for fbatch, bbatch in zip(fb, bb):
            outs = self.forward(blobs=blobs, **fbatch)
            dd = self.backward(diffs=diffs, **bbatch)
            for out, arr in outs.items():
                all_outs[out].extend(np.atleast_1d(arr).copy())
            for d, arr in dd.items():
                if d in all_diffs:
                    all_diffs[d].extend(arr.copy())
        for out in all_outs:
            all_outs[out] = np.asarray(all_outs[out])[:n]
        for d in all_diffs:
            all_diffs[d] = np.asarray(all_diffs[d])[:n]
        return all_outs, all_diffs

    def set_input_arrays(self, data, labels):
        """pycaffe.py:243-254 (MemoryData helper): fill the first two net inputs"""
        if labels.ndim == 1:
            labels = np.ascontiguousarray(labels[:, np.newaxis, np.newaxis, np.newaxis])
        names = self.inputs
        self.blobs[names[0]].data[...] = data
        self.blobs[names[1]].data[...] = labels.reshape(self.blobs[names[1]].shape)

    # device-resident training state (extensions used by the solver / the gradient all-reduce)
    def arenas(self):
        """(param_ptr, grad_ptr, count): fp32 device arenas holding every parameter blob / its gradient in layer order"""
        p, g, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        check(lib().eco_net_param_arena(self._h, C.byref(p), C.byref(n)))
        check(lib().eco_net_grad_arena(self._h, C.byref(g), C.byref(n)))
        return p.value, g.value, n.value

    def param_slots(self):
        n = C.c_int()
        check(lib().eco_net_num_param_slots(self._h, C.byref(n)))
        out = []
        for i in range(n.value):
            sl = _caffe.ParamSlot()
            check(lib().eco_net_param_slot(self._h, i, C.byref(sl)))
            out.append(dict(layer=self._layer_names[sl.layer], blob=sl.blob, offset=sl.offset, count=sl.count,
                            lr_mult=sl.lr_mult, decay_mult=sl.decay_mult))
        return out

    def params_updated_on_device(self):
        check(lib().eco_net_params_updated_on_device(self._h))

    def cuda_stream(self):
        s = C.c_void_p()
        check(lib().eco_net_cuda_stream(self._h, C.byref(s)))
        return s.value or 0

    def forward_all(self, blobs=None, **kwargs):
        all_outs = {out: [] for out in set(self.outputs + list(blobs or []))}
        n = len(next(iter(kwargs.values())))
        for batch in self._batch(kwargs):
            outs = self.forward(blobs=blobs, **batch)
            for out, arr in outs.items():
                all_outs[out].extend(arr.copy())
        for out in all_outs:
            all_outs[out] = np.asarray(all_outs[out])[:n]
        return all_outs

    def _batch(self, blobs):
        num = len(next(iter(blobs.values())))
        batch_size = self.blobs[self.inputs[0]].shape[0]
        for i in range(0, num - num % batch_size, batch_size):
            yield {name: blobs[name][i:i + batch_size] for name in blobs}
        rem = num % batch_size
        if rem:
            padded = {}
            for name in blobs:
                pad = np.zeros((batch_size - rem,) + blobs[name].shape[1:], np.float32)
                padded[name] = np.concatenate([blobs[name][-rem:], pad])
            yield padded

    def reshape(self):
        check(lib().eco_net_reshape(self._h))

    def sync(self):
        check(lib().eco_net_sync(self._h))

    # ---- weights --------------------------------------------------------------------
    def copy_from(self, path):
        check(lib().eco_net_copy_from(self._h, os.fsencode(path)))

    def save(self, path):
        check(lib().eco_net_save(self._h, os.fsencode(path)))

    def share_with(self, other):
        """caffe shares the parameter storage (net.cpp ShareTrainedLayersWith); here the values are
        copied once, matched by layer name."""
        src = other.params
        for name, blobs in self.params.items():
            if name in src:
                for a, b in zip(blobs, src[name]):
                    a.data[...] = b.data

    # ---- extensions beyond pycaffe ---------------------------------------------------
    def set_option(self, key, value):
        check(lib().eco_net_set_option(self._h, key.encode(), int(value)))

    def set_stream(self, cuda_stream_ptr):
        check(lib().eco_net_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr))))

    def set_input_device(self, blob_name, dev_ptr, count):
        check(lib().eco_net_set_input_device(self._h, self._blob_names.index(blob_name), C.c_void_p(int(dev_ptr)),
                                             int(count)))

    def blob_device_ptr(self, blob_name):
        p = C.c_void_p()
        n = C.c_size_t()
        check(lib().eco_blob_device_f32(self._h, self._blob_names.index(blob_name), C.byref(p), C.byref(n)))
        return p.value, n.value

    def forward_pipelined(self, host_in_ptr, count, host_out_ptr, out_count):
        """Serving extension: enqueue a forward fed from (pinned) host memory, results to host memory;
        returns a ticket for wait().  The H2D copy of call k+1 overlaps the compute of call k."""
        t = C.c_int()
        check(lib().eco_net_forward_pipelined(self._h, C.c_void_p(int(host_in_ptr)), int(count),
                                              C.c_void_p(int(host_out_ptr)), int(out_count), C.byref(t)))
        return t.value

    def forward_pipelined_u8(self, host_in_ptr, count, mean, host_out_ptr, out_count):
        """As forward_pipelined, fed from raw uint8 frames [F,3,H,W]; `mean` (per channel) is subtracted on the GPU."""
        t = C.c_int()
        m = (C.c_float * len(mean))(*[float(v) for v in mean])
        check(lib().eco_net_forward_pipelined_u8(self._h, C.c_void_p(int(host_in_ptr)), int(count), m, len(mean),
                                                 C.c_void_p(int(host_out_ptr)), int(out_count), C.byref(t)))
        return t.value

    def wait(self, ticket):
        check(lib().eco_net_wait(self._h, int(ticket)))

    def last_launch_count(self):
        n = C.c_int()
        check(lib().eco_net_last_launch_count(self._h, C.byref(n)))
        return n.value

    def describe_plan(self):
        """one dict per planned op: what the planner chose (kernel, tile, grid) for the current shapes"""
        need = C.c_size_t()
        check(lib().eco_net_describe_plan(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib().eco_net_describe_plan(self._h, buf, need.value, C.byref(need)))
        rows = []
        for line in buf.value.decode().splitlines():
            parts = line.split(" ")
            d = {"name": parts[0]}
            for kv in parts[1:]:
                k, _, v = kv.partition("=")
                d[k] = int(v) if v.lstrip("-").isdigit() else v
            rows.append(d)
        return rows

    def profile_train(self, cap=4096):
        arr = (_caffe.OpTime * cap)()
        n = C.c_int()
        check(lib().eco_net_profile_train(self._h, arr, cap, C.byref(n)))
        return [dict(name=arr[i].name.decode(), kind=arr[i].kind, ms=arr[i].ms) for i in range(n.value)]

    def profile_forward(self, cap=1024):
        arr = (_caffe.OpTime * cap)()
        n = C.c_int()
        check(lib().eco_net_profile_forward(self._h, arr, cap, C.byref(n)))
        return [dict(name=arr[i].name.decode(), kind=arr[i].kind, ms=arr[i].ms, flops=arr[i].flops,
                     bytes=arr[i].bytes) for i in range(n.value)]


def set_mode_gpu():
    check(lib().eco_set_mode(1))


def set_mode_cpu():
    # accepted like caffe.set_mode_cpu(); a forward in this mode raises: there is no CPU path
    check(lib().eco_set_mode(0))


def set_device(i):
    check(lib().eco_set_device(int(i)))


def set_logging_disabled():
    pass


def device_count():
    n = C.c_int()
    check(lib().eco_device_count(C.byref(n)))
    return n.value
