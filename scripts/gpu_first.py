"""First GPU contact: parity of the CUDA path against the oracle and the reference decoder, and
kernel timings of both on the headline configurations.  Prints a report; exits non-zero on mismatch."""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

spec = importlib.util.spec_from_file_location("vitb200", os.path.join(ROOT, "gpu-accelerated-viterbi-decoder_b200", "__init__.py"))
V = importlib.util.module_from_spec(spec)
spec.loader.exec_module(V)


def parity():
    bad = 0
    n = 6400 * 32 + 64 + 32 * 1234
    for it in range(5):
        for mt in (0x00, 0x10, 0x20):
            for ot in (0x000, 0x100):
                opt = it | mt | ot
                if not V.options_valid(opt):
                    continue
                dec = V.ViterbiCUDA(opt)
                for sigma, zero in ((0.9, False), (0.0, True)):
                    nn = n if it < 3 else 64 + 32 * 3000
                    bits, packed, N = O.make_channel_det(nn, it, seed=31 + it, sigma=sigma, zero=zero)
                    mine = dec.run(packed, N)
                    orc = O.decode(opt, packed, N)
                    ok_o = np.array_equal(mine, orc)
                    msg = "opt=%#06x sigma=%.1f zero=%d words=%d mine==oracle:%s" % (opt, sigma, zero, orc.size, ok_o)
                    if V.options_valid_ref(opt) and O.ref_lib() is not None:
                        ref, _ = O.ref_decode(opt, packed, N)
                        ov = O.overrun_words(opt, N).astype(np.int64)
                        mask = np.ones(ref.size, bool); mask[ov] = False
                        ok_r = np.array_equal(mine[mask], ref[mask])
                        orc_ov = O.decode(opt, packed, N, flags=O.FLAG_REF_OVERRUN)
                        msg += " mine==ref(owned):%s ref==oracle_overrun_emu:%s ovwords=%d" % (ok_r, np.array_equal(ref, orc_ov), ov.size)
                        bad += (not ok_r)
                    bad += (not ok_o)
                    print(msg, flush=True)
                dec.close()
    return bad


def timing():
    rows = []
    for name, opt, n in (("hard_b32_o32_32M", 0x000, 32_000_000), ("s4_b16_o32_32M", 0x011, 32_000_000),
                         ("s8_b16_o16_64M", 0x112, 64_000_000), ("f_b32_o32_32M", 0x004, 32_000_000),
                         ("s4_f16_o16_32M", 0x121, 32_000_000), ("s16_b32_o32_32M", 0x003, 32_000_000),
                         ("h_b16_o32_32M", 0x010, 32_000_000)):
        it = opt & 0xF
        bits, packed, N = O.make_channel(n, it, snr_db=5.5, seed=5, prbs=True)
        M = O.message_len(opt, N)
        dec = V.ViterbiCUDA(opt, N)
        best = 1e9
        for i in range(6):
            out, ms = dec.run(packed, N, want_kernel_time=True)
            best = min(best, ms)
        errs = O.count_errors(opt, out, M, bits)
        row = "%-18s mine: %.3f ms = %.1f Gb/s  BER=%.2e  info=%s" % (name, best, M / best / 1e6, errs / M, dec.kernel_info())
        if V.options_valid_ref(opt) and O.ref_lib() is not None:
            rbest = 1e9
            for i in range(4):
                rout, rms = O.ref_decode(opt, packed, N)
                rbest = min(rbest, rms)
            ov = O.overrun_words(opt, N).astype(np.int64)
            mask = np.ones(rout.size, bool); mask[ov] = False
            row += "  | ref: %.3f ms = %.1f Gb/s  speedup %.2fx  equal(owned)=%s" % (rbest, M / rbest / 1e6, rbest / best, np.array_equal(out[mask], rout[mask]))
        print(row, flush=True)
        dec.close()


if __name__ == "__main__":
    t = time.time()
    bad = parity()
    print("parity mismatching cases:", bad, "(%.1fs)" % (time.time() - t), flush=True)
    timing()
    sys.exit(1 if bad else 0)
