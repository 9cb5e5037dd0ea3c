"""Shared helpers for the GPU parity tests: build a net from prototxt text through the product's
Python shim (-> C ABI -> CUDA), load the oracle's harness weights into it, and compare blobs.

Tolerance (written here once, used by every parity test): the device path stores bf16 and
accumulates in fp32; against the oracle run in its bf16-mirror mode (same rounding points,
DESIGN.md "Rounding contract") a blob must satisfy
    max|a-b| <= 1e-3 * max|b|   for single fused ops fed identical inputs, and
    ||a-b||_2 <= TOL_NET * ||b||_2 for whole networks (rounding-boundary flips accumulate).
"""
import numpy as np

TOL_OP = 1e-3
TOL_NET = 2e-3


def make_net(text, keep_all=True, a_mode=None, graph=False):
    import caffe
    opts = {"keep_all_blobs": 1 if keep_all else 0, "use_graph": 1 if graph else 0}
    if a_mode is not None:
        opts["a_mode"] = a_mode
    return caffe.Net.from_string(text, caffe.TEST, **opts)


def load_params(net, params):
    """params: {layer: [arrays]} from oracle.refnet.RefNet.params_dict()."""
    P = net.params
    for name, arrs in params.items():
        assert name in P, name
        assert len(P[name]) == len(arrs), (name, len(P[name]), len(arrs))
        for blob, a in zip(P[name], arrs):
            assert tuple(blob.shape) == tuple(a.shape), (name, blob.shape, a.shape)
            blob.data[...] = a


def rel_max(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


def describe_mismatch(a, b, name=""):
    d = np.abs(a.astype(np.float64) - b)
    idx = np.unravel_index(np.argmax(d), d.shape)
    bad = d > 1e-2 * max(np.abs(b).max(), 1e-30)
    return "%s shape=%s rel_max=%.3e rel_l2=%.3e worst@%s got=%.5f want=%.5f bad_frac=%.4f" % (
        name, a.shape, rel_max(a, b), rel_l2(a, b), idx, a[idx], b[idx], bad.mean())
