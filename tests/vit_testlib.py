"""Shared helpers for the test-suite (imported by path: the name `tests` is not a unique package)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "gpu-accelerated-viterbi-decoder_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")

ALL_OPTS = [it | mt | ot for it in range(5) for mt in (0x00, 0x10, 0x20) for ot in (0x000, 0x100)
            if not (mt == 0x10 and it == 3)]


def load_pkg():
    """The product package (directory name has hyphens, so it is loaded by path)."""
    name = "gpu_accelerated_viterbi_decoder_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


# generator-polynomial pairs the emulator is also built for (tests/test_code_parameters.py)
ALT_EMU_PAIRS = [(0o117, 0o155), (0o133, 0o171)]
ALT_POLYS = (0o117, 0o155)
ALT_LIB = os.path.join(PKG_DIR, "libvitb200_p%o_%o.so" % ALT_POLYS)


def load_pkg_variant(lib_path, name):
    """A second instance of the product package bound to another build of the library (VIT_B200_LIB is read when the
    package is imported), e.g. one compiled for other generator polynomials."""
    if name in sys.modules:
        return sys.modules[name]
    old = os.environ.get("VIT_B200_LIB")
    os.environ["VIT_B200_LIB"] = lib_path
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        if old is None:
            del os.environ["VIT_B200_LIB"]
        else:
            os.environ["VIT_B200_LIB"] = old
    return mod


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def owned_mask(O, opt, N, nwords):
    """False on the words the reference leaves to its O_B16 over-run store race."""
    import numpy as np
    m = np.ones(nwords, bool)
    m[O.overrun_words(opt, N).astype(np.int64)] = False
    if (opt & 0x100) and nwords:
        # 16-bit packs, odd last segment: the reference runs 16-32 stages past the end of its input buffer
        # (viterbi.cu:186,199-206), so the stream's final word depends on whatever follows enc_d in device memory
        # (profiles/r1_parity_fuzz.txt); our decoder treats bytes past the input as zeros.
        q, r = divmod(nwords, 6400)
        if ((q + (1 if 6399 < r else 0)) * 16) % 32 == 16:
            m[nwords - 1] = False
    return m
