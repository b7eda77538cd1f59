"""Thin helper for the GPU tests: a libsmcb200 handle plus torch tensors, calling the C-ABI directly."""
import ctypes as C

import numpy as np
import torch

import smcb200

_lib = smcb200._lib


class Abi:
    def __init__(self, n_max, d_max=8, device=0):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        rc = self.lib.smcb_create(device, C.byref(self.h))
        if rc != 0:
            raise RuntimeError(self.lib.smcb_last_error(None).decode())
        self.dev = torch.device("cuda", device)
        self.ck(self.lib.smcb_reserve(self.h, n_max, d_max))

    def ck(self, rc):
        _lib.check(self.h, rc)

    def close(self):
        if self.h:
            self.lib.smcb_destroy(self.h)
            self.h = None

    def t(self, a, dtype=torch.float64):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.dev).contiguous()

    def zeros(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype, device=self.dev)

    # -------------------------------------------------------------------------------------------
    def loglik(self, model, theta_rows, active=None):
        """theta_rows: [n, d] host array (reference layout); returns lk[n] (numpy)."""
        th = self.t(np.asarray(theta_rows, dtype=np.float64).T)     # SoA [d, n]
        d, n = th.shape
        lk = self.zeros(n)
        act = self.t(active, torch.uint8) if active is not None else None
        self.ck(self.lib.smcb_loglik(self.h, model, th.data_ptr(), n, n, d, act.data_ptr() if act is not None else None,
                                     lk.data_ptr(), None))
        torch.cuda.synchronize()
        return lk.cpu().numpy()

    def stats(self):
        out = np.zeros(24, dtype=np.int64)
        self.ck(self.lib.smcb_loglik_stats(self.h, out.ctypes.data))
        return out
