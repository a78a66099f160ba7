"""ctypes loader for the C golden model (oracle/libvitoracle.so) and, when built, the reference's own
CUDA decoder (oracle/_ref/libvitref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

HARD, SOFT4, SOFT8, SOFT16, FP32 = 0, 1, 2, 3, 4
M_B32, M_B16, M_FP16 = 0x00, 0x10, 0x20
O_B32, O_B16 = 0x000, 0x100
REG, DPX = 0x0000, 0x1000
DPX_TIES = 0x2000   # extension: the tie rule of the reference's (never instantiated) DPX code paths, see vit_oracle.h
FLAG_REF_OVERRUN = 1
SEGMENTS = 6400
EXTRA_L = 26


def build(force=False):
    so = os.path.join(_HERE, "libvitoracle.so")
    src = os.path.join(_HERE, "vit_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libvitoracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        sz = C.c_size_t
        for name in ("vo_input_size", "vo_message_len", "vo_output_size"):
            f = getattr(L, name)
            f.restype, f.argtypes = sz, [C.c_int, sz]
        L.vo_options_valid_ref.restype, L.vo_options_valid_ref.argtypes = C.c_int, [C.c_int]
        L.vo_decode.restype = C.c_int
        L.vo_decode.argtypes = [C.c_int, C.c_void_p, C.c_void_p, sz, C.c_int, C.c_int]
        L.vo_decode_segments.restype = C.c_int
        L.vo_decode_segments.argtypes = [C.c_int, C.c_void_p, C.c_void_p, sz, sz, sz, C.c_int, C.c_int]
        L.vo_segment_window.restype = None
        L.vo_segment_window.argtypes = [C.c_int, sz, sz, sz] + [C.POINTER(sz)] * 4
        L.vo_decode_window.restype = C.c_int
        L.vo_decode_window.argtypes = [C.c_int, C.c_void_p, sz, sz, C.c_void_p, sz, sz, sz, sz, C.c_int]
        L.vo_overrun_words.restype = sz
        L.vo_overrun_words.argtypes = [C.c_int, sz, C.c_void_p, sz]
        L.vo_encode.restype, L.vo_encode.argtypes = None, [C.c_void_p, sz, C.c_void_p]
        L.vo_pack.restype, L.vo_pack.argtypes = None, [C.c_int, C.c_void_p, sz, C.c_float, C.c_void_p]
        L.vo_prbs31.restype, L.vo_prbs31.argtypes = None, [C.c_uint32, C.c_void_p, sz]
        L.vo_count_errors.restype = C.c_uint64
        L.vo_count_errors.argtypes = [C.c_int, C.c_void_p, sz, C.c_void_p]
        L.vo_num_threads.restype = C.c_int
        L.vo_set_segments.restype, L.vo_set_segments.argtypes = None, [sz]
        L.vo_set_polynomials.restype, L.vo_set_polynomials.argtypes = C.c_int, [C.c_uint, C.c_uint]
        L.vo_get_polynomials.restype, L.vo_get_polynomials.argtypes = None, [C.POINTER(C.c_uint)] * 2
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def input_size(options, n):
    return lib().vo_input_size(options, n)


def message_len(options, n):
    return lib().vo_message_len(options, n)


def output_size(options, n):
    return lib().vo_output_size(options, n)


def out_dtype(options):
    return np.uint16 if (options & 0xF00) == O_B16 else np.uint32


def decode(options, packed, input_num, nthreads=0, flags=0, segs=None):
    """packed: contiguous numpy array holding vo_input_size(options, input_num) bytes."""
    packed = np.ascontiguousarray(packed)
    assert packed.nbytes >= input_size(options, input_num), (packed.nbytes, input_size(options, input_num))
    out = np.zeros(output_size(options, input_num) // np.dtype(out_dtype(options)).itemsize, out_dtype(options))
    if segs is None:
        rc = lib().vo_decode(options, _ptr(packed), _ptr(out), input_num, nthreads, flags)
    else:
        rc = lib().vo_decode_segments(options, _ptr(packed), _ptr(out), input_num, segs[0], segs[1], nthreads, flags)
    if rc != 0:
        raise ValueError("oracle does not define option combination 0x%x" % options)
    return out


def segment_window(options, input_num, seg_begin, seg_end):
    """(in_byte0, in_bytes, out_word0, out_words) of segments [seg_begin, seg_end): the input bytes they read and the
    decoded packs they own."""
    v = [C.c_size_t(0) for _ in range(4)]
    lib().vo_segment_window(options, input_num, seg_begin, seg_end, *[C.byref(x) for x in v])
    return tuple(int(x.value) for x in v)


def decode_window(options, in_window, input_num, seg_begin, seg_end, nthreads=0):
    """Decode segments [seg_begin, seg_end) of a stream of which only the byte range segment_window() reports is given
    (in_window).  Returns (out_word0, packs) -- for streams too long to hold on the host."""
    b0, nb, w0, nw = segment_window(options, input_num, seg_begin, seg_end)
    in_window = np.ascontiguousarray(in_window)
    assert in_window.nbytes >= nb, (in_window.nbytes, nb)
    out = np.zeros(nw, out_dtype(options))
    rc = lib().vo_decode_window(options, _ptr(in_window), b0, nb, _ptr(out), w0, input_num, seg_begin, seg_end, nthreads)
    if rc != 0:
        raise ValueError("oracle does not define option combination 0x%x" % options)
    return w0, out


def decode_chunked(options, packed, chunk_syms, flags=0):
    """Restatement of the chunked stream decode (C ABI vit_stream_push): window k = symbols carried from window k-1 ++
    chunk k, decoded like a one-shot run() of that window; carry = the window from stage M_k on.  packed: the whole
    stream's channel words; chunk_syms: coded symbols per push (each a whole number of 32-bit packs).
    Returns (list of per-push outputs, pending symbols)."""
    it = options & 0xF
    spw = {HARD: 32, SOFT4: 8, SOFT8: 4, SOFT16: 2, FP32: 1}[it]
    words = np.ascontiguousarray(packed).view(np.uint32)
    bpp = 16 if (options & 0xF00) == O_B16 else 32
    outs, start, fed = [], 0, 0                 # start: first symbol of the current window
    for c in chunk_syms:
        assert c % spw == 0
        fed += c
        n = fed - start
        M = message_len(options, n)
        w = words[start // spw: fed // spw]
        outs.append(decode(options, w, n, flags=flags) if M else np.zeros(0, out_dtype(options)))
        start += 2 * M
        assert M % bpp == 0
    return outs, fed - start


def overrun_words(options, input_num):
    n = lib().vo_overrun_words(options, input_num, None, 0)
    idx = np.zeros(max(n, 1), np.uint64)
    lib().vo_overrun_words(options, input_num, _ptr(idx), n)
    return idx[:n]


def encode(bits):
    bits = np.ascontiguousarray(bits, np.uint8)
    coded = np.empty(2 * bits.size, np.uint8)
    lib().vo_encode(_ptr(bits), bits.size, _ptr(coded))
    return coded


def pack(input_type, soft, scale=1.0):
    """soft: float32 array of channel values (+-1 based).  Returns the packed encPack_t stream."""
    soft = np.ascontiguousarray(soft, np.float32)
    n = soft.size
    if input_type == FP32:
        out = np.empty(n, np.float32)
    else:
        per = {HARD: 32, SOFT4: 8, SOFT8: 4, SOFT16: 2}[input_type]
        if n % per:  # fill the last int32 pack with zero-valued symbols
            soft = np.concatenate([soft, np.zeros(per - n % per, np.float32)])
            n = soft.size
        out = np.empty(n // per, np.int32)
    lib().vo_pack(input_type, _ptr(soft), n, scale, _ptr(out))
    return out


def prbs31(seed, n):
    bits = np.empty(n, np.uint8)
    lib().vo_prbs31(seed, _ptr(bits), n)
    return bits


def count_errors(options, out, message_len_, bits):
    out = np.ascontiguousarray(out)
    bits = np.ascontiguousarray(bits, np.uint8)
    return int(lib().vo_count_errors(options, _ptr(out), message_len_, _ptr(bits)))


def set_segments(w):
    """test hook: segment count used by decode/overrun_words (0 -> the reference's 6400)."""
    lib().vo_set_segments(w)


def set_polynomials(polyn1=0, polyn2=0):
    """test hook: generator polynomials of the decoder model and of encode() (0, 0 -> the reference's 0171, 0133)."""
    if lib().vo_set_polynomials(polyn1, polyn2) != 0:
        raise ValueError("polynomials must be 7-bit values tapping bits 0 and 6 (got 0%o, 0%o)" % (polyn1, polyn2))


def get_polynomials():
    a, b = C.c_uint(0), C.c_uint(0)
    lib().vo_get_polynomials(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def num_threads():
    return lib().vo_num_threads()


def make_channel(n_bits, input_type, snr_db=None, seed=1, scale=40000.0, prbs=False, sigma=None):
    """Seeded twin of the reference harness's source -> encoder -> AWGN -> packer chain
    (main.cpp:131-138): sigma = 10^(-snr/5), scale 40000.  Returns (bits, packed, input_num)."""
    rng = np.random.default_rng(seed)
    bits = prbs31(0x7FFFFFFF ^ seed, n_bits) if prbs else rng.integers(0, 2, n_bits, dtype=np.uint8)
    coded = encode(bits)
    soft = coded.astype(np.float32) * 2.0 - 1.0
    if sigma is None and snr_db is not None:
        sigma = 10.0 ** (-snr_db / 5.0)
    if sigma:
        soft = soft + rng.standard_normal(soft.size, dtype=np.float32) * np.float32(sigma)
    return bits, pack(input_type, soft, scale), 2 * n_bits


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def hash_bits(seed, n_bits):
    """Message bits as a counter hash: bit i = top bit of splitmix64(i + seed * 0x9E3779B97F4A7C15 ... ).
    Twin of the device generator (csrc/vit_synth.cu) -- any bit can be produced independently."""
    with np.errstate(over="ignore"):
        h = _splitmix64(np.arange(n_bits, dtype=np.uint64) + np.uint64(seed) * np.uint64(0xD1B54A32D192ED03))
    return (h >> np.uint64(63)).astype(np.uint8)


def make_channel_det(n_bits, input_type, seed=1, amp=None, sigma=0.0, zero=False, bits_source="prbs"):
    """Bit-reproducible channel (integer arithmetic only, no libm, no numpy RNG): PRBS-31 message,
    K=7 encoder, symbol = +-amp + noise in Q8 fixed point, noise = (sum of four 16-bit uniforms from
    splitmix64, centred) scaled to a standard deviation of about sigma*amp, then the reference's
    quantise/clamp and MSB-first packing (viterbiDF.h:105-166).  Used for the golden vectors.
    Returns (bits, packed, input_num)."""
    if amp is None:
        amp = {HARD: 64, SOFT4: 3, SOFT8: 40, SOFT16: 9000, FP32: 48}[input_type]
    bits = prbs31(0x7FFFFFFF ^ seed, n_bits) if bits_source == "prbs" else hash_bits(seed, n_bits)
    coded = encode(bits).astype(np.int64)
    per = {HARD: 32, SOFT4: 8, SOFT8: 4, SOFT16: 2, FP32: 1}[input_type]
    nsym = coded.size
    v = (2 * coded - 1) * (amp << 8)
    sigma_q16 = int(round(sigma * amp * 256 / 0.57735))
    if sigma_q16:
        with np.errstate(over="ignore"):
            h = _splitmix64(np.arange(nsym, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x100000001B3))
        u = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)) +
             ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + (h >> np.uint64(48))).astype(np.int64) - 2 * 65535
        v = v + ((u * sigma_q16) >> 16)
    v = v >> 8
    if zero:
        v = np.zeros_like(v)
    if input_type == FP32:
        return bits, (v.astype(np.float32) / np.float32(16.0)), 2 * n_bits
    if nsym % per:
        v = np.concatenate([v, np.zeros(per - nsym % per, np.int64)])
    if input_type == HARD:
        q = (v > 0).astype(np.uint32)
        width = 1
    else:
        width = {SOFT4: 4, SOFT8: 8, SOFT16: 16}[input_type]
        lo, hi = -(1 << (width - 1)), (1 << (width - 1)) - 1
        q = (np.clip(v, lo, hi) & ((1 << width) - 1)).astype(np.uint32)
    q = q.reshape(-1, per)
    word = np.zeros(q.shape[0], np.uint32)
    for j in range(per):
        word = (word << np.uint32(width)) | q[:, j] if width < 32 else q[:, j]
    return bits, word.view(np.int32), 2 * n_bits


# ---------------------------------------------------------------------------------------------
# the reference's own CUDA decoder (needs a GPU to run; size helpers work anywhere)

_ref = None


def ref_lib():
    """oracle/_ref/libvitref.so or None."""
    global _ref
    if _ref is None:
        so = os.path.join(_HERE, "_ref", "libvitref.so")
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        L.ref_run.restype = C.c_int
        L.ref_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_float)]
        L.ref_sizes.restype = C.c_int
        L.ref_sizes.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ref_device_count.restype = C.c_int
        _ref = L
    return _ref


_ref_dpx = None


def ref_dpx_lib():
    """oracle/_ref/libvitref_dpx.so (the reference sources with compMode forwarded to forwardACS, ref_dpx_shim.cu) or None."""
    global _ref_dpx
    if _ref_dpx is None:
        so = os.path.join(_HERE, "_ref", "libvitref_dpx.so")
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        L.ref_run_dpx.restype = C.c_int
        L.ref_run_dpx.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_float)]
        _ref_dpx = L
    return _ref_dpx


def ref_decode_dpx(options, packed, input_num):
    """Run the reference decoder with its DPX code paths live (GPU).  options: the reference's own bitfield with
    CompMode DPX (0x1000).  Returns (out, kernel_ms)."""
    packed = np.ascontiguousarray(packed)
    nwords = output_size(options, input_num) // np.dtype(out_dtype(options)).itemsize
    out = np.zeros(nwords + 8, out_dtype(options))
    ms = C.c_float(0)
    rc = ref_dpx_lib().ref_run_dpx(options, _ptr(packed), _ptr(out), input_num, C.byref(ms))
    if rc != 0:
        raise ValueError("reference rejects option combination 0x%x" % options)
    return out[:nwords], ms.value


def ref_sizes(options, n):
    v = (C.c_size_t * 3)()
    rc = ref_lib().ref_sizes(options, n, v)
    return None if rc else (v[0], v[1], v[2])


def ref_decode(options, packed, input_num):
    """Run the unmodified reference decoder (GPU).  Returns (out, kernel_ms)."""
    packed = np.ascontiguousarray(packed)
    nwords = output_size(options, input_num) // np.dtype(out_dtype(options)).itemsize
    out = np.zeros(nwords + 8, out_dtype(options))  # slack: the reference's O_B16 over-run stores are device-side only
    ms = C.c_float(0)
    rc = ref_lib().ref_run(options, _ptr(packed), _ptr(out), input_num, C.byref(ms))
    if rc != 0:
        raise ValueError("reference rejects option combination 0x%x" % options)
    return out[:nwords], ms.value
