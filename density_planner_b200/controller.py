"""Controller weights: packing of the C3M controller into the blob the C ABI takes.

The trained controller is the pickled `U_FUNC` module of the reference
(data/trained_controller/model_CAR.py:9-28, loaded by systems/sytem_CAR.py:113-125).  Its parameters
(two MLPs 4-128-128-{48,24}, tanh, with bias; 43,592 fp32 values) are kernel constants here.
"""
import io
import os
import pickle
import sys
import types

import numpy as np
import torch

from . import _lib

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "controller_CAR_3layers.npz")
_KEYS = [("model_u_w1", 48), ("model_u_w2", 24)]


def pack_state_dict(sd):
    """state_dict of U_FUNC -> fp32 blob in the layout of include/density_planner_b200.h (dp_controller_create)."""
    parts = []
    for net, nout in _KEYS:
        shapes = {"0.weight": (128, 4), "0.bias": (128,), "2.weight": (128, 128), "2.bias": (128,),
                  "4.weight": (nout, 128), "4.bias": (nout,)}
        for layer in ("0", "2", "4"):
            for kind in ("weight", "bias"):
                key = "%s.%s.%s" % (net, layer, kind)
                v = sd[key]
                v = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
                if tuple(v.shape) != shapes["%s.%s" % (layer, kind)]:
                    raise _lib.DensityPlannerError("controller tensor %s has shape %s, expected %s (only the "
                                                   "3-layer CAR controller is supported)" % (key, v.shape, shapes[layer + "." + kind]))
                parts.append(np.ascontiguousarray(v, dtype=np.float32).ravel())
    blob = np.concatenate(parts)
    assert blob.size == _lib.DP_CONTROLLER_BLOB_FLOATS
    return blob


def _load_pickled_module(path):
    """Load the reference's pickled U_FUNC without needing its model_CAR.py on sys.path."""

    class U_FUNC(torch.nn.Module):  # attribute container only; forward is never called
        pass

    shim = types.ModuleType("model_CAR")
    shim.U_FUNC = U_FUNC
    had = sys.modules.get("model_CAR")
    sys.modules["model_CAR"] = shim
    try:
        obj = torch.load(path, map_location="cpu", weights_only=False)
    finally:
        if had is None:
            sys.modules.pop("model_CAR", None)
        else:
            sys.modules["model_CAR"] = had
    return obj


def load_blob(path=None):
    """Packed controller weights from (a) the packaged .npz, (b) another .npz, (c) the reference's .pth.tar pickle."""
    if path is None:
        path = _DATA
    if not os.path.exists(path):
        raise _lib.DensityPlannerError("controller weights not found at %s" % path)
    if path.endswith(".npz"):
        z = np.load(path)
        return pack_state_dict({k: z[k] for k in z.files})
    obj = _load_pickled_module(path)
    sd = obj.state_dict() if hasattr(obj, "state_dict") else obj
    return pack_state_dict(sd)


def random_blob(seed=0):
    """Random-init weights of the same architecture (torch.nn.Linear default init) for synthetic benchmarks."""
    g = torch.Generator().manual_seed(seed)
    parts = []
    for _, nout in _KEYS:
        for fan_in, fan_out in ((4, 128), (128, 128), (128, nout)):
            bound = 1.0 / np.sqrt(fan_in)
            parts.append(((torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound).numpy().ravel())
            parts.append(((torch.rand(fan_out, generator=g) * 2 - 1) * bound).numpy().ravel())
    return np.concatenate(parts).astype(np.float32)


class DeviceController:
    """Owns a dp_controller handle (device copies of the weights).  The handle is created on first use so that the
    host-side logic (sampling, reference integration) can be constructed and tested on a box without a GPU; any
    compute call without a B200 fails loudly."""

    def __init__(self, blob, device):
        self.blob = np.ascontiguousarray(blob, np.float32)
        self._device = device
        self._handle = None

    def _create(self):
        import ctypes
        L = _lib.lib()
        if not torch.cuda.is_available():
            raise _lib.DensityPlannerError("density_planner_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(self._device if self._device is not None else "cuda")
        if dev.type != "cuda":
            raise _lib.DensityPlannerError("density_planner_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self._device = torch.device("cuda", idx)
        h = ctypes.c_void_p()
        _lib.check(L.dp_controller_create(self.blob.ctypes.data_as(ctypes.c_void_p), self.blob.size, idx, ctypes.byref(h)),
                   "dp_controller_create")
        self._handle = h

    @property
    def handle(self):
        if self._handle is None:
            self._create()
        return self._handle

    @property
    def device(self):
        if self._handle is None:
            self._create()
        return self._device

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().dp_controller_destroy(self._handle)
                self._handle = None
        except Exception:
            pass
