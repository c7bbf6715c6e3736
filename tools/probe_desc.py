"""Does a UMMA smem descriptor tolerate a start row that is not a multiple of 8 (tap sharing of one halo tile)?"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from promptttspp_b200 import _abi  # noqa: E402

lib = _abi.lib()
g = torch.Generator().manual_seed(0)
A = (torch.randn(160, 64, generator=g)).half().cuda()
Bm = (torch.randn(128, 64, generator=g)).half().cuda()
out = torch.empty(128, 128, device="cuda")
for mode in (0, 1):
    for off in (0, 1, 2, 3, 4, 7, 8, 9, 15, 16):
        out.fill_(float("nan"))
        _abi.check(lib.pttspp_umma_probe(_abi.ptr(A), 160, _abi.ptr(Bm), off, mode, _abi.ptr(out), _abi.stream_ptr()))
        torch.cuda.synchronize()
        ref = A[off:off + 128].float() @ Bm.float().t()
        err = float((out - ref).abs().max())
        print(f"mode {mode} row_off {off:2d}: max-abs err {err:.3e} {'OK' if err < 1e-2 else 'MISMATCH'}", flush=True)
