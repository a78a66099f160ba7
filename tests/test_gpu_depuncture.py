"""GPU test (-m gpu) of the depuncturing pre-pass (vit_depuncture_device): punctured soft streams of the K=7 (171,133)
mother code at the standard rates are expanded to rate 1/2 with erasures and decoded by the ordinary kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WIDTH = {1: 4, 2: 8, 3: 16}


def _pack(it, vals):
    """MSB-first packing of integer symbols into int32 words (reference viterbiDF.h:139-166); FP32: floats."""
    if it == 4:
        return vals.astype(np.float32)
    w = WIDTH[it]
    per = 32 // w
    pad = (-vals.size) % per
    v = np.concatenate([vals, np.zeros(pad, np.int64)]) & ((1 << w) - 1)
    v = v.reshape(-1, per).astype(np.uint32)
    word = np.zeros(v.shape[0], np.uint32)
    for j in range(per):
        word = (word << np.uint32(w)) | v[:, j]
    return word


@pytest.mark.parametrize("it", [1, 2, 3, 4])
@pytest.mark.parametrize("rate", ["1/2", "2/3", "3/4", "5/6", "7/8"])
def test_depuncture_matches_numpy_twin_and_decodes(V, O, it, rate):
    import torch
    n_bits = 200_000 + 7
    period, k0, k1 = V.PUNCTURE[rate]
    bits = O.prbs31(11 + it, n_bits)
    coded = O.encode(bits).astype(np.int64)                        # 2 symbols per stage, 0171 first
    amp = {1: 7, 2: 100, 3: 20000, 4: 6}[it]
    rng = np.random.default_rng(5)
    lo, hi = (-(1 << (WIDTH[it] - 1)), (1 << (WIDTH[it] - 1)) - 1) if it != 4 else (-8, 7)
    sym = np.clip((2 * coded - 1) * amp + np.rint(rng.normal(0, amp * 0.2, coded.size)).astype(np.int64), lo, hi)
    stage = np.arange(coded.size) // 2
    which = np.arange(coded.size) % 2
    keep = np.where(which == 0, (k0 >> (stage % period)) & 1, (k1 >> (stage % period)) & 1).astype(bool)
    tx = sym[keep]                                                  # what the transmitter sends
    expect = np.where(keep, sym, 0)                                 # numpy twin of the expansion
    d_tx = torch.from_numpy(_pack(it, tx).view(np.uint8).copy()).cuda()
    nbytes = _pack(it, expect).nbytes
    d_full = torch.zeros(nbytes + 64, dtype=torch.uint8, device="cuda")
    V.depuncture_device(it, d_tx.data_ptr(), tx.size, rate, d_full.data_ptr(), n_bits)
    torch.cuda.synchronize()
    assert np.array_equal(d_full[:nbytes].cpu().numpy(), _pack(it, expect).view(np.uint8)), (it, rate)
    # ... and the expanded stream decodes: no bit errors at this noise level, bit-exact with the golden model
    opt = it | 0x10 if it != 3 else it                              # int16x2 core (int32 for 16-bit symbols)
    dec = V.ViterbiCUDA(opt)
    N = 2 * n_bits
    d_out = torch.zeros(dec.getOutputSize(N) + 64, dtype=torch.uint8, device="cuda")
    dec.run_device(d_full.data_ptr(), d_out.data_ptr(), N)
    torch.cuda.synchronize()
    got = d_out[:dec.getOutputSize(N)].cpu().numpy().view(np.uint32)
    assert np.array_equal(got, O.decode(opt, _pack(it, expect), N))
    errs = O.count_errors(opt, got, dec.getMessageLen(N), bits)
    # The reference's window (32-stage warm-up, 38-stage traceback merge, viterbi.h:70-76) is sized for the rate-1/2 code;
    # punctured codes need deeper windows as the rate grows, so the higher rates leave a truncation error floor even at
    # this mild noise level.  Bit-exactness with the golden model above is the parity statement; here: sanity only.
    print("rate %s, input type %d: %d bit errors of %d" % (rate, it, errs, dec.getMessageLen(N)))
    if rate in ("1/2", "2/3"):
        assert errs == 0, (it, rate)
    else:
        assert errs < 0.05 * dec.getMessageLen(N), (it, rate, errs)
    dec.close()
