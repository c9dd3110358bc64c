"""CPU tests of bench.py's metric arithmetic: algorithmic FLOPs per refined chunk and per tensor-core launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench
from vla_touch_b200 import native as nv


def test_flops_per_chunk_matches_the_survey_figures():
    # cfg2: 2 x 12.25 (DinoV2-S @224) + 20 x 1.297 (U-Net, T=64) = 50.4 GFLOP; cfg3: DinoV2-B, 50 steps = 222.3 GFLOP
    name, hidden, heads, layers, hw, T, A, F, steps, batch = bench.WORKLOADS["cfg2"]
    assert abs(bench.flops_per_chunk(hidden, layers, hw, T, A, F, steps) / 1e9 - 50.4) < 0.1
    name, hidden, heads, layers, hw, T, A, F, steps, batch = bench.WORKLOADS["cfg3"]
    assert abs(bench.flops_per_chunk(hidden, layers, hw, T, A, F, steps) / 1e9 - 222.3) < 0.5


def test_launch_flops_of_gemm_and_fused_mlp_descriptors():
    g = nv.GemmDesc()
    g.G, g.M, g.N, g.taps, g.kc, g.passes = 2, 1024, 256, 5, 256, 1
    assert bench.gemm_flops(g) == 2.0 * 2 * 1024 * 256 * 5 * 256
    m = nv.MlpDesc()
    m.rows, m.D = 131584, 384
    assert bench.gemm_flops(m) == 2 * (2.0 * 131584 * 384 * 1536)      # fc1 + fc2
