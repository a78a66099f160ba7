"""GPU tests of the device-resident entry points (vit_run_device / vit_run_device_batch) and of the
full-size BASELINE configurations through size-independent properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_api_with_torch_buffers(V, O):
    import torch
    opt = 0x011
    bits, packed, N = O.make_channel_det(6400 * 32 * 3 + 64, O.SOFT4, seed=3, sigma=0.9)
    dec = V.ViterbiCUDA(opt)
    d_in = torch.from_numpy(packed.view(np.uint8).copy()).cuda()
    d_out = torch.zeros(dec.getOutputSize(N), dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        dec.run_device(d_in.data_ptr(), d_out.data_ptr(), N, stream=st.cuda_stream)
    st.synchronize()
    got = d_out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, O.decode(opt, packed, N))
    # misaligned device pointer is refused, not silently mis-read
    d_in2 = torch.zeros(d_in.numel() + 16, dtype=torch.uint8, device="cuda")
    with pytest.raises(V.ViterbiError):
        dec.run_device(d_in2.data_ptr() + 4, d_out.data_ptr(), N)
    dec.close()


def test_batched_streams(V, O):
    """Independent codeword streams in one launch (BASELINE config 5 shape, scaled down)."""
    import torch
    opt = 0x112
    ns, n = 5, 64 + 16 * 6400 * 3 + 16 * 5
    packs = [O.make_channel_det(n, O.SOFT8, seed=50 + s, sigma=0.9) for s in range(ns)]
    N = packs[0][2]
    dec = V.ViterbiCUDA(opt)
    in_stride = (dec.getInputSize(N) + 255) // 256 * 256
    out_stride = (dec.getOutputSize(N) + 255) // 256 * 256
    d_in = torch.zeros(ns * in_stride, dtype=torch.uint8, device="cuda")
    for s, (_, p, _) in enumerate(packs):
        d_in[s * in_stride: s * in_stride + p.nbytes] = torch.from_numpy(p.view(np.uint8).copy()).cuda()
    d_out = torch.zeros(ns * out_stride, dtype=torch.uint8, device="cuda")
    dec.run_device(d_in.data_ptr(), d_out.data_ptr(), N, nstreams=ns, in_stride=in_stride, out_stride=out_stride)
    torch.cuda.synchronize()
    host = d_out.cpu().numpy()
    for s, (_, p, _) in enumerate(packs):
        got = host[s * out_stride: s * out_stride + dec.getOutputSize(N)].view(np.uint16)
        assert np.array_equal(got, O.decode(opt, p, N)), s
    dec.close()


@pytest.mark.parametrize("opt,n,snr", [(0x011, 32_000_000, 15.0), (0x000, 32_000_000, 15.0), (0x000, 1_000_000, 5.5)])
def test_baseline_configs_full_size(V, O, opt, n, snr):
    """BASELINE configs 0-1 at full size: encode -> channel -> decode round trip is error free, and the
    whole output equals the golden model's (the oracle decodes 32 Mbit in a few seconds)."""
    bits, packed, N = O.make_channel(n, opt & 0xF, snr_db=snr, seed=21, prbs=True)
    dec = V.ViterbiCUDA(opt, N)
    out = dec.run(packed, N)
    M = dec.getMessageLen(N)
    assert O.count_errors(opt, out, M, bits) == 0
    assert np.array_equal(out, O.decode(opt, packed, N))
    dec.close()


def test_config3_f16_256M_mismatch_vs_b32(V, O):
    """BASELINE config 2: s8 input, half2 core, 16-bit packs, 256 Mbit at 3 dB; mismatch count against
    the int32 core on the same bytes is reported and must be 0 at this operating point."""
    import torch
    n = 256_000_000
    # build the channel on the device (plumbing): PRBS-free random bits, K=7 encoder, AWGN, s8 saturating quantiser
    g = torch.Generator(device="cuda").manual_seed(3)
    bits = torch.randint(0, 2, (n,), dtype=torch.int8, device="cuda", generator=g)
    pad = torch.zeros(6, dtype=torch.int8, device="cuda")
    b = torch.cat([pad, bits])
    def par(taps):
        acc = torch.zeros(n, dtype=torch.int8, device="cuda")
        for t in taps:
            acc ^= b[t: t + n]                                 # buffer bit t (6 = newest) at step i is bits[i-6+t]
        return acc
    o0, o1 = par([0, 3, 4, 5, 6]), par([0, 1, 3, 4, 6])       # 0171, 0133 (viterbiDF.h:48-60)
    del b
    sym = torch.stack([o0, o1], 1).reshape(-1).float() * 2 - 1
    del o0, o1
    sym += torch.randn(sym.shape, device="cuda", generator=g) * (10 ** (-3.0 / 5.0))
    q = torch.clamp(torch.round(sym * 40000.0), -128, 127).to(torch.int8)
    del sym
    # MSB-first packing into int32 = big-endian byte order within each 4-byte word
    d_in = q.view(-1, 4).flip(1).contiguous().view(torch.uint8).view(-1)
    N = 2 * n
    outs = {}
    for opt in (0x122, 0x102):
        dec = V.ViterbiCUDA(opt)
        d_out = torch.zeros(dec.getOutputSize(N), dtype=torch.uint8, device="cuda")
        dec.run_device(d_in.data_ptr(), d_out.data_ptr(), N)
        torch.cuda.synchronize()
        outs[opt] = d_out
        dec.close()
    mism = int((outs[0x122] != outs[0x102]).sum().item())
    assert mism == 0
    # round trip: out bit j == message bit j+26 (main.cpp:153-169); check on the device
    M = V.lib().vit_message_len(0x122, N)
    words = outs[0x122].view(torch.int16).view(-1)
    # u16 words, MSB = earliest
    w = words.to(torch.int32) & 0xFFFF
    shifts = torch.arange(15, -1, -1, device="cuda", dtype=torch.int32)
    dec_bits = ((w.unsqueeze(1) >> shifts) & 1).reshape(-1).to(torch.int8)
    errs = int((dec_bits[:M] != bits[26:26 + M]).sum().item())
    assert errs <= M * 1e-6, errs
    # a slice of the output is also checked against the golden model on the same bytes
    host_in = d_in.cpu().numpy()
    seg_out = O.decode(0x122, host_in, N, segs=(3000, 3016))
    P = M // 16
    qq, rr = divmod(P, 6400)
    a, bnd = qq * 3000 + min(3000, rr), qq * 3016 + min(3016, rr)
    assert np.array_equal(words.cpu().numpy().view(np.uint16)[a:bnd], seg_out[a:bnd])
