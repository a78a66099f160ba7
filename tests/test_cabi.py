"""CPU tests of the drop-in boundary: the C-ABI library loads, exports everything the header
declares, and its pure helpers agree with the reference's formulas.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

from vit_testlib import PKG_DIR, ROOT

HEADER = os.path.join(ROOT, "include", "vit_b200.h")
LIB = os.path.join(PKG_DIR, "libvitb200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vit_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("vit_create", "vit_destroy", "vit_run", "vit_run_device", "vit_run_device_batch",
              "vit_input_size", "vit_message_len", "vit_output_size", "vit_options_valid", "vit_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build it: python -c 'import __graft_entry__ as g; g.build()'"
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    # and nothing but the C ABI leaks out
    assert all(s.startswith("vit_") for s in exported), sorted(s for s in exported if not s.startswith("vit_"))[:5]


def test_library_contains_sm100a_code():
    out = subprocess.run(["cuobjdump", "--list-elf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_size_helpers_and_option_table(V, O):
    L = V.lib()
    for opt in range(0, 0x4000):
        it, mt, ot, cm = opt & 0xF, (opt >> 4) & 0xF, (opt >> 8) & 0xF, (opt >> 12) & 0xF
        known = it <= 4 and mt <= 2 and ot <= 1 and cm <= 2
        if not known:
            assert not L.vit_options_valid(opt)
            continue
        assert bool(L.vit_options_valid_ref(opt)) == bool(O.lib().vo_options_valid_ref(opt))
        # ours: superset (f16 x s8/s16 allowed, dpx = reg, plus the DPX tie rule as CompMode 2 for the integer cores);
        # b16 x s16 rejected as in the reference
        assert bool(L.vit_options_valid(opt)) == (not (mt == 1 and it == 3) and not (cm == 2 and mt == 2))
        for n in (0, 100, 128, 2_000_000, 12_345_679):
            assert L.vit_input_size(opt, n) == O.input_size(opt & 0xFFF, n)
            assert L.vit_message_len(opt, n) == O.message_len(opt & 0xFFF, n)
            assert L.vit_output_size(opt, n) == O.output_size(opt & 0xFFF, n)


def test_errors_are_reported_not_fatal(V):
    L = V.lib()
    h = C.c_void_p()
    assert L.vit_create(C.byref(h), 0x013, 0, 0) == 1          # VIT_ERR_OPTIONS: b16 x s16
    assert b"unsupported" in L.vit_last_error()
    with pytest.raises(V.ViterbiError):
        V.ViterbiCUDA(0x7)                                       # unknown input type


def test_parse_options_matches_main_flags(V):
    # reference main.cpp:211-254
    assert V.parse_options() == 0
    assert V.parse_options("s4", "b16", "b32") == 0x011
    assert V.parse_options("SOFT8", "f16", "b16") == 0x122
    assert V.parse_options("f", "b32", "b32", "dpx") == 0x1004
    assert V.parse_options("h", "b32", "b32", "dpxt") == 0x2000


def test_no_cpu_fallback_in_product_sources():
    """The product path must not reach into oracle/ or the emulator."""
    for dirpath, _, files in os.walk(PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("vit_oracle.h", "libvitoracle", "libvitref", "from oracle", "import oracle", "oracle/",
                               "vo_decode", "libvitemu", "vit_emu"):
                    assert needle not in txt, (dirpath, f, needle)


def test_one_lane_per_segment_geometry_cross_compiles(tmp_path):
    """csrc/vit_kernel_l1.inc (the many-stream geometry the library does not instantiate yet) builds for sm_100a: no
    spills for the packed cores, and its trellis stages contain no shuffle at all."""
    import shutil
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on this machine")
    obj = tmp_path / "l1_probe.o"
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-Xptxas", "-v",
                        "-I", os.path.join(PKG_DIR, "csrc"), "-c", os.path.join(ROOT, "scripts", "l1_probe.cu"), "-o", str(obj)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    blocks = re.split(r"Compiling entry function '", r.stderr)[1:]
    assert len(blocks) == 6
    for b in blocks:
        name = b.split("'")[0]
        regs = int(re.search(r"Used (\d+) registers", b).group(1))
        assert regs <= 200, (name, regs)
        if "ILi1E" in name or "ILi2E" in name:                       # int16x2 and half2 cores
            assert "0 bytes spill stores" in b, name
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    assert "SHFL" not in sass and "LDGSTS" in sass and "VIMNMX" in sass


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md maps every C entry point to the reference interface it replaces (families by prefix)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    families = ("vit_comm_", "vit_job_", "vit_dev_", "vit_host_", "vit_stream_")
    missing = [s for s in declared_symbols() if s not in doc and not s.startswith(families)]
    assert not missing, missing
    for fam in families:
        assert fam in doc, fam


def test_shipped_kernels_have_no_spills_and_the_expected_instructions():
    """static facts of the built library (csrc/build/*.ptxas.log from the build, cuobjdump of the objects): all 76 decode
    kernels without local memory, sm_100a SASS with the instructions the design rests on and none it must not have."""
    build = os.path.join(PKG_DIR, "csrc", "build")
    if not os.path.isdir(build):
        pytest.skip("library not built here")
    kernels = l1_kernels = 0
    for unit in ("b16", "b32", "b32d", "f16"):
        log = open(os.path.join(build, "vit_inst_%s.ptxas.log" % unit)).read()
        for block in re.split(r"Compiling entry function '", log)[1:]:
            name = block.split("'")[0]
            if "vit_decode_kernel" not in name:
                continue
            assert "0 bytes stack frame, 0 bytes spill stores, 0 bytes spill loads" in block, block[:200]
            regs = int(re.search(r"Used (\d+) registers", block).group(1))
            if "vit_decode_kernel_l1" in name:                             # the opt-in one-lane-per-segment geometry (packed cores)
                l1_kernels += 1
                assert regs <= 168
            else:
                kernels += 1
                assert regs <= 104                                         # >= 19 resident warps/SM by registers
    assert kernels == 76                                                   # (28 + 10 option combinations) x 2 table builds
    assert l1_kernels == 18                                                # b16: 4 input types x 2 pack widths, f16: 5 x 2
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(build, "vit_inst_b16.o")], capture_output=True, text=True).stdout
    for needle in ("VIMNMX.S16x2", "LDGSTS.E.BYPASS.128", "SHFL.BFLY", "STS.128", "LDS.64"):
        assert needle in sass, needle
    for absent in ("HMMA", "IMMA", "UTCHMMA", "UTMALDG", "LDL", "STL"):   # no tensor-core / TMA path by design, no local memory
        assert absent not in sass, absent
    sass32 = subprocess.run(["cuobjdump", "-sass", os.path.join(build, "vit_inst_b32.o")], capture_output=True, text=True).stdout
    assert "VIADDMNMX" in sass32
