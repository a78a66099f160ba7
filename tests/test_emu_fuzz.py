"""CPU fuzz of the product kernel SOURCE (host emulator, tests/emu) against the golden model: random option values,
stream lengths from below one pack to dozens of packs per segment, random segment counts, noise from none to heavy,
saturated and all-tie inputs, both operand-table builds, all four lane geometries (8 = the product's, 4, 16, and one lane per segment), staged and direct output stores.
Fixed seed, bounded time (15 s, several hundred cases).  (The same comparison on the sm_100a binary: scripts/parity_fuzz.py inside the -m gpu suite.)"""
import time

import numpy as np

from test_emu_kernel import emu_decode
from vit_testlib import ALL_OPTS

BUDGET_S = 15.0
OPTS = ALL_OPTS + [0x2000 | o for o in ALL_OPTS if (o & 0xF0) != 0x20]      # + the DPX tie rule for the integer cores


def test_kernel_source_fuzz(emu, O):
    rng = np.random.default_rng(20261018)
    t0, cases, by_geom = time.perf_counter(), 0, {8: 0, 4: 0, 16: 0, 1: 0}
    try:
        while time.perf_counter() - t0 < BUDGET_S:
            opt = int(OPTS[rng.integers(len(OPTS))])
            it, bpp = opt & 0xF, (16 if opt & 0x100 else 32)
            W = int(rng.integers(1, 33))
            packs = int(rng.integers(0, W * 12)) if rng.random() < 0.8 else int(rng.integers(0, 4))
            n = 64 + packs * bpp + int(rng.integers(0, bpp))                 # ragged tail below one pack
            kind = rng.random()
            kw = dict(seed=int(rng.integers(1, 1 << 30)))
            if kind < 0.15:
                kw["zero"] = True                                            # every compare a tie
            elif kind < 0.35:
                kw.update(sigma=float(rng.uniform(0.0, 0.4)), amp={0: 64, 1: 7, 2: 127, 3: 32767, 4: 128}[it])   # saturating
            else:
                kw["sigma"] = float(rng.choice([0.0, 0.3, 0.8, 1.5, 3.0]))
            zero = kw.pop("zero", False)
            bits, packed, N = O.make_channel_det(max(n, 64), it, zero=zero, **kw)
            lanes = int(rng.choice([8, 8, 4, 16, 1]))
            tbl = int(rng.choice([96, 32]))
            staged = int(rng.random() < 0.3)
            emu.vit_emu_set_lanes(lanes); emu.vit_emu_set_table(tbl); emu.vit_emu_set_stage_out(staged)
            O.set_segments(W)
            ref = O.decode(opt, packed, N)
            got = emu_decode(emu, O, opt, packed, N, W)
            assert np.array_equal(got, ref), dict(opt=hex(opt), n=n, W=W, lanes=lanes, tbl=tbl, staged=staged, **kw, zero=zero)
            cases += 1
            by_geom[lanes] += 1
    finally:
        O.set_segments(0)
        emu.vit_emu_set_lanes(8); emu.vit_emu_set_table(96); emu.vit_emu_set_stage_out(0)
    assert cases >= 50 and all(v > 0 for v in by_geom.values()), (cases, by_geom)
