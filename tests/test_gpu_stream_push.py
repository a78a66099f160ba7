"""GPU tests (-m gpu) of the chunked stream decode (C ABI vit_stream_push): an endless stream pushed in chunks of arbitrary
(whole-pack) sizes.  Parity definition (include/vit_b200.h): window k = symbols carried from window k-1 ++ chunk k is
decoded exactly like a one-shot run() of that window -- checked against the golden model's restatement
(oracle.decode_chunked) and against the LIVE reference decoder run on the same windows -- and the concatenated outputs
are the contiguous message bits 26, 27, ... of the stream."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SPW = {0: 32, 1: 8, 2: 4, 3: 2, 4: 1}


def _chunks(rng, total_syms, spw, lo, hi):
    out, left = [], total_syms
    while left >= spw:
        c = min(int(rng.integers(lo, hi)) * spw, left // spw * spw)
        out.append(c)
        left -= c
    return out


@pytest.mark.parametrize("opt", [0x011, 0x000, 0x112, 0x004, 0x121, 0x023, 0x2001])
def test_stream_push_equals_oracle_and_reference_windows(V, O, opt):
    it = opt & 0xF
    spw = SPW[it]
    n_bits = 3_000_000 if it < 3 else 400_000
    bits, packed, N = O.make_channel_det(n_bits, it, seed=90 + it, sigma=0.8)
    rng = np.random.default_rng(opt)
    for lo, hi in ((1, 40), (2000, 60000)):               # tiny chunks (many below one window's 64 stages), then large ones
        total = N if hi > 1000 else min(N, 40000 * spw)
        chunks = _chunks(rng, total, spw, lo, hi)
        exp, pend = O.decode_chunked(opt, packed, chunks)
        dec = V.ViterbiCUDA(opt)
        dec.stream_reset()
        words = np.ascontiguousarray(packed).view(np.uint32)
        pos, got = 0, []
        for c in chunks:
            got.append(dec.stream_push(words[pos // spw:(pos + c) // spw], c).copy())
            pos += c
        assert dec.stream_pending() == pend
        for k, (g, e) in enumerate(zip(got, exp)):
            assert np.array_equal(g, e), (hex(opt), k, chunks[k])
        allw = np.concatenate(got)
        bpp = dec.bitsPerPack
        assert dec.stream_bits() == allw.size * bpp
        # the concatenation is the contiguous message: same error positions as ... at least as many bits as a one-shot decode minus one pack per push
        assert allw.size * bpp >= O.message_len(opt, sum(chunks)) - bpp * len(chunks)
        # live reference on the same windows (only combinations the reference accepts; large chunks only)
        if hi > 1000 and O.ref_lib() is not None and V.options_valid_ref(opt):
            start = fed = 0
            for k, c in enumerate(chunks):
                fed += c
                n = fed - start
                M = O.message_len(opt, n)
                if M:
                    ref, _ = O.ref_decode(opt, words[start // spw: fed // spw], n)
                    ov = O.overrun_words(opt, n).astype(np.int64)
                    m = np.ones(ref.size, bool)
                    m[ov] = False
                    if (opt & 0x100) and ref.size:
                        m[-1] = False                      # the reference's odd last segment reads past its input (vit_testlib.owned_mask)
                    assert np.array_equal(got[k][m], ref[m]), (hex(opt), k)
                start += 2 * M
        dec.close()


def test_stream_push_noiseless_is_the_contiguous_message(V, O):
    opt, n_bits = 0x012, 2_000_000
    bits, packed, N = O.make_channel_det(n_bits, O.SOFT8, seed=5, sigma=0.0)
    dec = V.ViterbiCUDA(opt)
    words = packed.view(np.uint32)
    outs, pos = [], 0
    rng = np.random.default_rng(3)
    for c in _chunks(rng, N, 4, 1, 30000):
        outs.append(dec.stream_push(words[pos // 4:(pos + c) // 4], c).copy())
        pos += c
    allw = np.concatenate(outs)
    M = allw.size * 32
    assert M == dec.stream_bits() and M >= n_bits - 64 - 32
    assert O.count_errors(opt, allw, M, bits) == 0
    # chunks must be whole 32-bit packs; the output buffer must hold what the chunk completes
    with pytest.raises(V.ViterbiError):
        dec.stream_push(words[:1], 3)
    dec.stream_reset()
    assert dec.stream_pending() == 0 and dec.stream_bits() == 0
    dec.close()


def test_stream_push_device_buffers(V, O):
    import ctypes as C
    import torch
    opt, n_bits = 0x011, 1_500_000
    bits, packed, N = O.make_channel_det(n_bits, O.SOFT4, seed=6, sigma=0.7)
    exp, pend = O.decode_chunked(opt, packed, [N // 3 // 8 * 8, N // 3 // 8 * 8, N // 3 // 8 * 8])
    dec = V.ViterbiCUDA(opt)
    d_in = torch.from_numpy(packed.view(np.uint8).copy()).cuda()
    d_out = torch.zeros(dec.getOutputSize(N) + 256, dtype=torch.uint8, device="cuda")
    c = N // 3 // 8 * 8
    off = 0
    for k in range(3):
        n = C.c_size_t(0)
        rc = V.lib().vit_stream_push_device(dec._h, d_in.data_ptr() + k * (c // 2), c, d_out.data_ptr() + off, d_out.numel() - off, C.byref(n), None)
        assert rc == 0
        torch.cuda.synchronize()
        got = d_out[off:off + n.value].cpu().numpy().view(np.uint32)
        assert np.array_equal(got, exp[k]), k
        off += n.value
    dec.close()
