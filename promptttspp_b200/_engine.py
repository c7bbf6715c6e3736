"""Shared plumbing between the nn.Module shims and the native handles.

A shim owns torch Parameters/buffers under the reference's key names.  The native handle keeps
its own packed copy of them; `NativeHandle.sync()` re-uploads whenever the parameters changed
(load_state_dict, .to(), remove_weight_norm_, in-place edits), detected through the tensors'
identity and version counters.
"""
import ctypes as C

import torch

from . import _abi


class NativeHandle:
    def __init__(self, kind, cfg_struct):
        self.kind = kind
        self._lib = _abi.lib()
        self._h = C.c_void_p()
        create = getattr(self._lib, f"pttspp_{kind}_create")
        _abi.check(create(C.byref(cfg_struct), C.byref(self._h)))
        self._sig = None
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                getattr(self._lib, f"pttspp_{self.kind}_destroy")(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @staticmethod
    def signature(tensors):
        return tuple((k, t.data_ptr(), t._version, tuple(t.shape), str(t.device)) for k, t in tensors.items())

    def sync(self, tensors, device):
        """tensors: {reference key: torch tensor}; uploads + finalizes if anything changed."""
        sig = self.signature(tensors)
        if sig == self._sig:
            return
        set_tensor = getattr(self._lib, f"pttspp_{self.kind}_set_tensor")
        stream = _abi.stream_ptr(device)
        for name, t in tensors.items():
            src = t.detach()
            if src.dtype != torch.float32:
                src = src.float()
            src = src.contiguous()
            shape = (C.c_int64 * max(src.dim(), 1))(*src.shape)
            _abi.check(set_tensor(self._h, name.encode(), _abi.ptr(src), shape, src.dim(), stream))
        _abi.check(getattr(self._lib, f"pttspp_{self.kind}_finalize")(self._h, stream))
        self._sig = sig

    def invalidate(self):
        """Forget the upload signature: the next sync() re-packs every tensor (for writes through `.data`, which do not
        bump the version counters the signature watches)."""
        self._sig = None

    def workspace(self, nbytes, device):
        """Grow-only scratch buffer from torch's caching allocator."""
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self._ws

    @property
    def h(self):
        return self._h
