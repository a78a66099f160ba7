"""CPU model check of vit_run's time-sliced upload (csrc/vit_api.cu run_gated + the gates in csrc/vit_kernel_map.inc).

One decode launch starts before its input has arrived; the copy stream uploads column block b of EVERY segment row and then
opens gate b; a warp passes gate g before it requests the channel words of super-step x >= gate_super[g].  The property
that makes this correct -- and that a GPU test can only sample, because a violation shows up as a timing-dependent wrong
word -- is checked here exhaustively over many geometries:

    every input byte a segment USES in super-step x has been enqueued for upload before the flag of the last gate g with
    gate_super[g] <= x.

The plan and the kernel's byte ranges are restated from the sources (formulas cited below); a guard test fails if the cited
expressions disappear from the sources, so that the model cannot silently drift from the code."""
import os
import re

import numpy as np
import pytest

from vit_testlib import PKG_DIR

W = 6400
B96 = {0: 24, 1: 96, 2: 192, 3: 384, 4: 768}          # channel bytes per 96 stages (vit_api.cu host_run_geometry, InTraits)
PER = {0: 32, 1: 8, 2: 4, 3: 2, 4: 1}


def geometry(opt, input_num):
    """vit_api.cu host_run_geometry"""
    it, bpp = opt & 0xF, (16 if opt & 0x100 else 32)
    in_bytes = {0: (input_num + 7) // 8, 1: (input_num + 1) // 2, 2: input_num, 3: 2 * input_num, 4: 4 * input_num}[it]
    M = (input_num // 2 - 64) // bpp * bpp if input_num // 2 >= 64 else 0
    P = M // bpp
    q, r = divmod(P, W)
    b96 = B96[it]
    pack_bytes = bpp * b96 // 96
    Lmax = (q + (1 if r else 0)) * bpp
    Tmax = 64 + 32 * ((Lmax + 31) // 32)
    return dict(it=it, bpp=bpp, in_bytes=in_bytes, P=P, q=q, r=r, b96=b96, pack_bytes=pack_bytes, nsuper=(Tmax + 95) // 96)


def gate_plan(g, staged):
    """vit_api.cu run_gated: gate_super[] of the pinned plan (3 blocks) and of the staged plan (equal blocks)"""
    ns = g["nsuper"]
    if staged:
        n = min(4, max(2, ns // 4))
        return [ns * b // n for b in range(n)]
    if g["it"] == 0:
        return [0, ns // 8, ns // 2]
    return [0, ns // 2, ns * 7 // 8]


def uploaded_columns(g, sup, b):
    """columns [0, hi) of every row that blocks 0..b have covered (block b: [super[b]*b96 - 16, super[b+1]*b96 + 48), the
    blocks overlap, so the union is contiguous from column 0); None = whole rows"""
    if b + 1 == len(sup):
        return None
    return sup[b + 1] * g["b96"] + 48


def check_geometry(opt, input_num, staged):
    g = geometry(opt, input_num)
    sup = gate_plan(g, staged)
    assert sup == sorted(sup) and sup[0] == 0
    q, r, bpp, b96 = g["q"], g["r"], g["bpp"], g["b96"]
    w = np.arange(W, dtype=np.int64)
    Lp = q + (w < r)                                                     # packs per segment (viterbi.cu:156-165)
    row = (q * w + np.minimum(w, r)) * g["pack_bytes"]                   # first byte of the segment = its row start
    pitch = Lp * g["pack_bytes"]
    body_end = g["P"] * g["pack_bytes"]
    T = 64 + 32 * ((Lp * bpp + 31) // 32)                                # stages whose results the segment uses
    first = w - w % 4                                                    # the warp's first segment sets its super-step count
    T_warp = 64 + 32 * (((q + (first < r)) * bpp + 31) // 32)
    nsuper_w = (T_warp + 95) // 96
    assert nsuper_w.max() == g["nsuper"]
    checked = 0
    for gi in range(len(sup)):
        hi = uploaded_columns(g, sup, gi)
        if hi is None:
            continue                                                     # after the last gate everything has been enqueued
        # super-steps released by gate gi: x < sup[gi + 1]; the bytes used up to there, relative to the row start
        stages = np.minimum(96 * sup[gi + 1], T)
        stages = np.where(sup[gi + 1] > nsuper_w, T, stages)            # (a short warp never gets that far)
        need_end = row + (stages * b96 + 95) // 96                       # exclusive, absolute
        need_end = np.minimum(need_end, g["in_bytes"])                   # bytes past the input are zero-filled, never read
        live = Lp > 0
        # walk the rows the needed range [row_w, need_end) crosses: row w' contributes [row', row' + min(hi, pitch'))
        cur = row.copy()
        wp = w.copy()
        for _ in range(8):
            covered_to = row[wp] + np.minimum(hi, pitch[wp])
            full_row = hi >= pitch[wp]
            nxt = np.where(full_row, row[wp] + pitch[wp], covered_to)
            cur = np.maximum(cur, np.minimum(nxt, need_end))
            done = (cur >= need_end) | ~full_row | (wp == W - 1)
            if np.all(done | ~live):
                break
            wp = np.where(done, wp, np.minimum(wp + 1, W - 1))
        # the stream's last 64 stages (beyond the last row) go with the first block
        cur = np.where((cur >= body_end), need_end, cur)
        bad = live & (cur < need_end)
        assert not bad.any(), dict(opt=hex(opt), input_num=input_num, staged=staged, gate=gi, plan=sup, segment=int(w[bad][0]),
                                   short_by=int((need_end - cur)[bad][0]), **{k: g[k] for k in ("q", "r", "nsuper", "b96")})
        checked += int(live.sum())
    return g, sup, checked


def applies(g, forced):
    """vit_api.cu gated_upload_applies"""
    if forced:
        return g["q"] >= 1 and g["nsuper"] >= 8
    return g["q"] >= 1 and g["nsuper"] >= 16 and g["in_bytes"] >= (2 << 20)


@pytest.mark.parametrize("opt", [0x000, 0x100, 0x011, 0x111, 0x012, 0x112, 0x023, 0x004, 0x104])
def test_upload_plan_covers_every_used_byte(opt):
    rng = np.random.default_rng(opt + 7)
    it, bpp = opt & 0xF, (16 if opt & 0x100 else 32)
    sizes = []
    for q in list(range(1, 60)) + [100, 156, 157, 333, 1000, 2499, 2500, 7812, 19531]:
        for r in (0, 1, 3, 4, 5, 1234, 3200, 6396, 6397, 6399):
            bits = (q * W + r) * bpp + 64 + int(rng.integers(0, bpp))
            sizes.append(2 * bits + int(rng.integers(0, 2)))
    sizes += [2 * 1_000_000, 2 * 32_000_000, 2 * 256_000_000]            # BASELINE sizes
    n_geom = n_checked = 0
    for input_num in sizes:
        input_num -= input_num % PER[it]                                 # whole 32-bit channel packs, as the callers pass
        g = geometry(opt, input_num)
        for forced in (False, True):
            if not applies(g, forced):
                continue
            for staged in (False, True):
                _, _, c = check_geometry(opt, input_num, staged)
                n_checked += c
                n_geom += 1
    assert n_geom > 200 and n_checked > 10**6, (n_geom, n_checked)


def test_the_model_would_catch_a_short_block():
    """the check is not vacuous: a plan whose first block stops 64 bytes early misses bytes of the super-step it releases"""
    global uploaded_columns
    keep = uploaded_columns
    try:
        uploaded_columns = lambda g, sup, b: None if b + 1 == len(sup) else sup[b + 1] * g["b96"] - 64
        with pytest.raises(AssertionError):
            check_geometry(0x011, 2 * 32_000_000, False)
    finally:
        uploaded_columns = keep


def test_model_formulas_are_the_ones_in_the_sources():
    api = open(os.path.join(PKG_DIR, "csrc", "vit_api.cu")).read()
    kmap = open(os.path.join(PKG_DIR, "csrc", "vit_kernel_map.inc")).read()
    flat = re.sub(r"\s+", " ", api)
    for needle in ("const size_t lo = b == 0 ? 0 : (size_t)gp.super[b] * g.b96 - 16;",
                   "const size_t hi = b + 1 == gp.n ? (size_t)-1 : (size_t)gp.super[b + 1] * g.b96 + 48;",
                   "gp.super[1] = (unsigned)(g.nsuper / 8); gp.super[2] = (unsigned)(g.nsuper / 2);",
                   "gp.super[1] = (unsigned)(g.nsuper / 2); gp.super[2] = (unsigned)(g.nsuper * 7 / 8);",
                   "gp.super[b] = (unsigned)(g.nsuper * b / gp.n);",
                   "std::min<size_t>(env_blocks > 0 ? (size_t)std::min(env_blocks, 8) : 4, std::max<size_t>(2, g.nsuper / 4))",
                   "if (g.q < 1 || g.nsuper < 16 || g.in_bytes < (2u << 20)) return false;",
                   "return h->gate_d && h->gate_err_d && g.q >= 1 && g.nsuper >= 8;",
                   "const size_t Lmax = (g.q + (g.r ? 1 : 0)) * g.bpp, Tmax = 64 + 32 * ((Lmax + 31) / 32);",
                   "if (g.in_bytes > body_end) {"):
        assert needle in flat, needle
    kflat = re.sub(r"\s+", " ", kmap)
    for needle in ("const unsigned Tmax = 64 + 32 * ((Lmax + 31) / 32);", "const unsigned nsuper = (Tmax + SUPER - 1) / SUPER;",
                   "VIT_GATE(sc + 1) \\ issue_raw_copy(c, sc + 1);", "VIT_GATE(0u) issue_raw_copy(c, 0);",
                   "while (gate_next < kp.gate_n && (x) >= gate_super_at(kp, gate_next))"):
        assert needle in kflat, needle


@pytest.mark.parametrize("opt", [0x000, 0x011, 0x112, 0x023, 0x004])
def test_chunk_pipeline_uploads_what_its_kernels_use(opt):
    """vit_api.cu run_chunked (pinned buffers where gates cannot be used): the stream is cut at segment boundaries
    (multiples of 8 segments = whole warps) into nch chunks; chunk i's kernel starts when the input bytes [0, in_hi_i) have
    landed.  Every byte the segments of chunk i use must lie below in_hi_i (or beyond the input: zero-filled)."""
    it, bpp = opt & 0xF, (16 if opt & 0x100 else 32)
    rng = np.random.default_rng(opt + 99)
    n_geom = 0
    for q in (40, 41, 97, 156, 157, 333, 1000, 2499, 2500, 7812):
        for r in (0, 1, 7, 8, 9, 1234, 3199, 3200, 3201, 6393, 6399):
            input_num = 2 * ((q * W + r) * bpp + 64 + int(rng.integers(0, bpp)))
            input_num -= input_num % PER[it]
            g = geometry(opt, input_num)
            nch = min(8, g["in_bytes"] // (4 << 20))
            if nch < 2:
                continue
            q_, r_, b96 = g["q"], g["r"], g["b96"]
            w = np.arange(W, dtype=np.int64)
            Lp = q_ + (w < r_)
            start_pack = q_ * w + np.minimum(w, r_)
            used_end = np.minimum(((start_pack * bpp + 64 + 32 * ((Lp * bpp + 31) // 32)) * b96 + 95) // 96, g["in_bytes"])
            for i in range(nch):
                a = W * i // nch // 8 * 8
                b = W if i + 1 == nch else W * (i + 1) // nch // 8 * 8
                if i + 1 == nch:
                    in_hi = g["in_bytes"]
                else:
                    last_bits = (q_ + (1 if b - 1 < r_ else 0)) * bpp
                    end_stage = (q_ * (b - 1) + min(b - 1, r_)) * bpp + 64 + 32 * ((last_bits + 31) // 32) + 32
                    in_hi = min(g["in_bytes"], (end_stage * b96 + 95) // 96 + 64)
                assert a % 4 == 0 and b % 4 == 0 and b > a
                assert used_end[a:b].max() <= in_hi, dict(opt=hex(opt), input_num=input_num, chunk=i, nch=nch)
            n_geom += 1
    assert n_geom > 50
    flat = re.sub(r"\s+", " ", open(os.path.join(PKG_DIR, "csrc", "vit_api.cu")).read())
    for needle in ("const size_t end_stage = g.start_pack(b - 1) * g.bpp + 64 + 32 * ((last_bits + 31) / 32) + 32;",
                   "in_hi = std::min(g.in_bytes, (end_stage * g.b96 + 95) / 96 + 64);",
                   "g.nch = (int)std::min<size_t>(vit_handle::MAX_CHUNKS, g.in_bytes / (4u << 20));",
                   "const unsigned a = (unsigned)(g.W * i / g.nch / 8 * 8)"):
        assert needle in flat, needle
