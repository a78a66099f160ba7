"""GPU test of the experimental one-lane-per-segment geometry through the library (vit_set_geometry(L1), csrc/vit_kernel_l1.inc).
NOT part of the default GPU suite: the kernel itself has run on a B200 (scripts/l1_bench.cu, profiles/r2_l1_bench.txt: outputs
identical to the 8-lane kernel), but this route through the library's launch table has only been exercised in the CPU
host-path simulation (tests/test_sim_python_mirror.py) -- there was no GPU time left in the round that wrote it.  Run it with
VIT_TEST_L1=1 before making L1 more than an opt-in."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("VIT_TEST_L1") != "1", reason="experimental geometry: set VIT_TEST_L1=1")]


@pytest.mark.parametrize("opt", [0x012, 0x011, 0x110, 0x024, 0x123])
def test_one_lane_geometry_equals_the_golden_model(V, O, opt):
    import torch
    ns, n = 13, 64 + 32 * 6400 * 3 + 32 * 5
    packs = [O.make_channel_det(n, opt & 0xF, seed=50 + s, sigma=0.9) for s in range(ns)]
    N = packs[0][2]
    dec = V.ViterbiCUDA(opt)
    dec.set_geometry(V.GEOMETRY_L1)
    in_stride = (dec.getInputSize(N) + 255) // 256 * 256
    out_stride = (dec.getOutputSize(N) + 255) // 256 * 256
    d_in = torch.zeros(ns * in_stride, dtype=torch.uint8, device="cuda")
    for s, (_, p, _) in enumerate(packs):
        raw = np.ascontiguousarray(p).view(np.uint8)[:dec.getInputSize(N)].copy()
        d_in[s * in_stride: s * in_stride + raw.size] = torch.from_numpy(raw).cuda()
    d_out = torch.zeros(ns * out_stride, dtype=torch.uint8, device="cuda")
    ms = dec.run_device(d_in.data_ptr(), d_out.data_ptr(), N, nstreams=ns, in_stride=in_stride, out_stride=out_stride, want_kernel_time=True)
    assert ms > 0 and dec.last_launch_geometry() == V.GEOMETRY_L1
    host = d_out.cpu().numpy()
    for s, (_, p, _) in enumerate(packs):
        got = host[s * out_stride: s * out_stride + dec.getOutputSize(N)].view(dec.decPack_t)
        assert np.array_equal(got, O.decode(opt, p, N)), s
    dec.close()
