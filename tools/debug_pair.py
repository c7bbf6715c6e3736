import math, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["PTTSPP_UMMA_PAIR"] = "2"
os.environ["PTTSPP_UMMA_DEBUG"] = "1"
from promptttspp_b200 import ops
import torch.nn.functional as F
torch.set_grad_enabled(False)
for (Cin, Cout, K, dil, B, T) in [(256, 512, 3, 1, 2, 300), (256, 256, 1, 1, 3, 77), (256, 512, 3, 8, 4, 2500), (256, 512, 3, 2, 16, 2582)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / math.sqrt(Cin * K)
    b = torch.randn(Cout, generator=g)
    pad = (K * dil - dil) // 2
    ref = F.conv1d(x, w, b, padding=pad, dilation=dil)
    planes = ops.split_f16(x.transpose(1, 2).contiguous().cuda())
    try:
        out, _ = ops.conv1d_umma_cl(planes, ops.pack_conv_weight_split(w, device="cuda"), Cout, bias=b.cuda(), K=K, dil=dil, pad=pad)
        torch.cuda.synchronize()
        err = float((out.float().cpu().transpose(1, 2) - ref).abs().max())
        print((Cin, Cout, K, dil, B, T), "max-abs err", err, flush=True)
    except Exception as e:
        print((Cin, Cout, K, dil, B, T), "FAILED", e, flush=True)
