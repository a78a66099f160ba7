"""GPU tests of vit_run's host-buffer paths (-m gpu): pinned / pageable / mixed buffers, every upload mode, and the
environments in which copies cannot run beside a kernel (where the time-sliced upload must not be attempted)."""
import os
import subprocess
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _buffers(torch, packed, in_bytes, out_bytes, pin_in, pin_out):
    raw = packed.view(np.uint8)[:in_bytes].copy()
    h_in = torch.from_numpy(raw).pin_memory().numpy() if pin_in else raw
    h_out_t = torch.zeros(out_bytes, dtype=torch.uint8)
    h_out = h_out_t.pin_memory().numpy() if pin_out else h_out_t.numpy()
    return h_in, h_out


@pytest.mark.parametrize("pin_in,pin_out", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("opt,n", [
    (0x011, 6400 * 32 * 52 + 32 * 1234 + 64 + 5),    # s4, ragged: the first 1234 segments are one pack longer
    (0x000, 6400 * 32 * 330 + 32 * 77 + 64),          # hard input: 24 channel bytes per super-step
    (0x112, 6400 * 16 * 101 + 16 * 3 + 64),           # 16-bit packs, odd pack count per segment
])
def test_host_run_pinned_pageable_mixes(V, O, opt, n, pin_in, pin_out):
    """Every mix of pinned and pageable caller buffers takes the time-sliced upload (ONE launch per run; pageable
    sides are staged by the worker threads) and decodes to the golden model's output.  A pinned input with a pageable
    output used to stall for the gate time-out (the pageable download blocked the host before the uploads were issued)."""
    import torch
    bits, packed, N = O.make_channel_det(n, opt & 0xF, seed=78, sigma=0.7)
    dec = V.ViterbiCUDA(opt, N)
    h_in, h_out = _buffers(torch, packed, dec.getInputSize(N), dec.getOutputSize(N), pin_in, pin_out)
    exp = O.decode(opt, packed, N)
    launches = dec.launch_count()
    t0 = time.perf_counter()
    for rep in range(3):
        h_out[:] = 0
        dec.run(h_in, N, output_h=h_out.view(dec.decPack_t))
        assert np.array_equal(h_out.view(dec.decPack_t), exp), rep
    assert time.perf_counter() - t0 < 1.0                     # no gate time-out anywhere
    assert dec.launch_count() - launches == 3
    assert dec.upload_mode_in_effect() == V.UPLOAD_GATED
    dec.close()


def test_upload_modes_give_identical_output(V, O):
    import torch
    opt, n = 0x011, 6400 * 32 * 55 + 32 * 17 + 64          # 11.3 MB of input: two chunks of the segment-range pipeline
    bits, packed, N = O.make_channel_det(n, opt & 0xF, seed=79, sigma=0.8)
    exp = O.decode(opt, packed, N)
    dec = V.ViterbiCUDA(opt, N)
    for pinned in (True, False):
        h_in, h_out = _buffers(torch, packed, dec.getInputSize(N), dec.getOutputSize(N), pinned, pinned)
        for mode, launches in ((V.UPLOAD_SEQUENTIAL, 1), (V.UPLOAD_CHUNKED, 2 if pinned else 1), (V.UPLOAD_GATED, 1), (V.UPLOAD_AUTO, 1)):
            dec.set_upload_mode(mode)
            before = dec.launch_count()
            h_out[:] = 0
            dec.run(h_in, N, output_h=h_out.view(dec.decPack_t))
            assert np.array_equal(h_out.view(dec.decPack_t), exp), (pinned, mode)
            assert dec.launch_count() - before == launches, (pinned, mode)
    with pytest.raises(V.ViterbiError):
        dec.set_upload_mode(7)
    # output_h is validated before its pointer reaches the library
    with pytest.raises(V.ViterbiError):
        dec.run(packed, N, output_h=np.zeros(10, dec.decPack_t))
    with pytest.raises(V.ViterbiError):
        dec.run(packed, N, output_h=np.zeros(2 * exp.size, dec.decPack_t)[::2])
    dec.close()


_ENV_CODE = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from vit_testlib import load_pkg
from oracle import oracle as O
V = load_pkg()
opt, n = 0x011, 6400 * 32 * 55 + 64
bits, packed, N = O.make_channel_det(n, opt & 0xF, seed=3, sigma=0.7)
exp = O.decode(opt, packed, N)
dec = V.ViterbiCUDA(opt, N)
assert dec.upload_mode_in_effect() == V.UPLOAD_CHUNKED, dec.upload_mode_in_effect()
raw = packed.view(np.uint8)[:dec.getInputSize(N)].copy()
for pinned in (True, False):
    h_in = torch.from_numpy(raw).pin_memory().numpy() if pinned else raw
    h_out = torch.zeros(dec.getOutputSize(N), dtype=torch.uint8)
    h_out = (h_out.pin_memory() if pinned else h_out).numpy()
    dec.run(h_in, N, output_h=h_out.view(dec.decPack_t))          # first call: context warm-up
    t0 = time.perf_counter()
    for rep in range(3):
        h_out[:] = 0
        dec.run(h_in, N, output_h=h_out.view(dec.decPack_t))
        assert np.array_equal(h_out.view(dec.decPack_t), exp), (pinned, rep)
    dt = time.perf_counter() - t0
    assert dt < 0.5, dt
print("OK", dec.launch_count())
'''


@pytest.mark.parametrize("env", [{"CUDA_DEVICE_MAX_CONNECTIONS": "1"}, {"CUDA_LAUNCH_BLOCKING": "1"}])
def test_environments_where_copies_cannot_overlap_a_kernel(env):
    """With one hardware queue for all streams, or with blocking launches, a copy issued after the decode kernel cannot
    run before it ends: vit_run must see that up front, take the chunk pipeline and never spend a gate time-out."""
    code = _ENV_CODE % (os.path.join(ROOT, "tests"), ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_device_calls_keep_the_callers_current_device(V, O):
    """API calls run on the handle's device and restore the thread's current device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    bits, packed, N = O.make_channel_det(64 + 32 * 6400 * 2, O.SOFT4, seed=5, sigma=0.5)
    torch.cuda.set_device(0)
    dec = V.ViterbiCUDA(0x011, N, device=1)
    out = dec.run(packed, N)
    assert torch.cuda.current_device() == 0
    assert np.array_equal(out, O.decode(0x011, packed, N))
    dec.close()
