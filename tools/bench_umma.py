"""Manual GPU tuning aid (not a test): cycles per tcgen05.mma for the shapes the conv kernels issue.

mode bits: 1 A from tensor memory (TS), 2 B MN-major, 4 M=64, 8 rotate 9 accumulators, 16 operands one row off the 128-byte
core-matrix alignment (a tap shift), 32 a tcgen05.cp of the next A tile every 18 MMAs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import howl_b200

ctx = howl_b200.Context(torch.device("cuda:0"))
names = {0: "SS  B K-major", 16: "SS  B K-major, misaligned", 2: "SS  B MN-major", 1: "TS  B K-major", 3: "TS  B MN-major",
         3 + 8: "TS  MN, 9 accumulators", 3 + 16: "TS  MN, misaligned", 3 + 32: "TS  MN, cp every 18", 3 + 8 + 16 + 32: "TS  MN, all three",
         4: "SS  M=64 B K-major"}
for mode, name in names.items():
    row = []
    for n in (16, 48, 64, 96, 128, 256):
        if (mode & 8) and n > 48:
            continue
        ctx.debug_umma_bench(mode, n, 72)
        row.append(f"N={n}: {ctx.debug_umma_bench(mode, n, 4608):6.1f}")
    print(f"{name:28s}", "  ".join(row), flush=True)
