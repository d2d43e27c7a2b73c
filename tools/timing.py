"""Bring-up helper: per-role barrier wait cycles of CTA 0 for the headline InfoNCE shape."""
import os, subprocess, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RANGECLIP_B200_LIB"] = os.path.abspath("rangeclip_b200/librangeclip_b200_timing.so")
from rangeclip_b200 import _lib, ops
import ctypes
_lib.lib().rc_debug_set_timing_buffer.argtypes = [ctypes.c_void_p]      # (a bring-up entry point: not in _lib.PROTOTYPES)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
D, H, W, K = 512, 256, 256, 256
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, D, H, W, device=dev, generator=g).to(torch.bfloat16)
t = torch.nn.functional.normalize(torch.randn(K, D, device=dev, generator=g), dim=1)
y = torch.randint(0, K, (B, H * W), device=dev, generator=g, dtype=torch.int32)
w = torch.ones(B, H * W, device=dev)
buf = torch.zeros(2, 48, device=dev, dtype=torch.int64)
roles = ["producer", "mma", "compute"]
names = {0: {0: "lifetime", 1: "empty(S)", 2: "empty(dX)"},
         1: {0: "lifetime", 3: "s_empty", 4: "full(S)", 5: "p_full", 6: "acc_empty", 7: "full(dX)"},
         2: {0: "lifetime", 8: "s_full", 9: "p_empty", 11: "acc_full", 12: "stg_full", 13: "named_bar", 14: "wait_read",
             15: "tmem_ld", 5: "smx_gload", 6: "smx_pass1", 7: "smx_pass2", 10: "smx_tail", 1: "epi_compute", 2: "fence.proxy", 3: "store+commit", 4: "issue_load"}}
for rep in range(2):
    _lib.lib().rc_debug_set_timing_buffer(buf[rep].data_ptr())
    r = ops.infonce_raw(x, t, y, w, 1 / 0.07, True, False, "bf16")
    torch.cuda.synchronize()
_lib.lib().rc_debug_set_timing_buffer(None)
tiles = (B * H * W // 128 + 147) // 148
row = buf[1].tolist()
for r_, role in enumerate(roles):
    vals = row[r_ * 16:(r_ + 1) * 16]
    print(f"{role}: " + ", ".join(f"{names[r_].get(i, i)}={v / tiles:.0f}" for i, v in enumerate(vals) if v))
