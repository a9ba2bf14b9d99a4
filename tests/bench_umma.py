"""Manual GPU tuning aid (not a test): cycles per tcgen05.mma for the shapes the conv kernels issue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import howl_b200

ctx = howl_b200.Context(torch.device("cuda:0"))
names = {0: "SS  B K-major", 2: "SS  B MN-major", 1: "TS  B K-major", 3: "TS  B MN-major", 4: "SS  M=64 B K-major", 7: "TS  M=64 B MN-major"}
for mode, name in names.items():
    row = []
    for n in (16, 32, 48, 64, 96, 128, 192, 256):
        ctx.debug_umma_bench(mode, n, 64)
        row.append(f"N={n}: {ctx.debug_umma_bench(mode, n, 4096):6.1f}")
    print(f"{name:22s}", "  ".join(row), flush=True)
