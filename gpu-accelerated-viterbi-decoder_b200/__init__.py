"""B200-native Viterbi decoder: Python mirror of the reference's host class.

`ViterbiCUDA(options)` mirrors `template<int options> class ViterbiCUDA`
(reference src/viterbi/viterbi.h:91-152): same method names, same argument meaning (every size is
a count of CODED SYMBOLS), same option bitfield.  Every call goes through the C ABI in
include/vit_b200.h (libvitb200.so, hand-written sm_100a kernels).  There is no CPU path: if the
shared library is missing or CUDA fails, these functions raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# VIT_B200_LIB: another build of the same library (e.g. one compiled for other generator polynomials, csrc/vit_code.h).  An
# explicit choice made before import, never a fallback: if the named file is missing, lib() raises like for the default.
LIB_PATH = os.environ.get("VIT_B200_LIB") or os.path.join(_HERE, "libvitb200.so")

# option bitfield, reference src/viterbi/viterbi.h:7-20
HARD, SOFT4, SOFT8, SOFT16, FP32 = 0x0, 0x1, 0x2, 0x3, 0x4
M_B32, M_B16, M_FP16 = 0x00, 0x10, 0x20
O_B32, O_B16 = 0x000, 0x100
REG, DPX = 0x0000, 0x1000
DPX_TIES = 0x2000   # extension (not a reference value): int32 core with the tie rule of the reference's dead DPX code paths
CHANNEL_MASK, METRIC_MASK, DECODE_MASK, COMP_MASK = 0xF, 0xF0, 0xF00, 0xF000

# window constants, reference viterbi.h:61-79
constLen, polyn1, polyn2 = 7, 0o171, 0o133
extraL, extraR, slideSize, forwardLen = 26, 38, 32, 96
SEGMENTS = 6400
UPLOAD_AUTO, UPLOAD_SEQUENTIAL, UPLOAD_CHUNKED, UPLOAD_GATED = 0, 1, 2, 3
GEOMETRY_L8, GEOMETRY_L1 = 0, 1          # vit_set_geometry: lane geometry of gate-free launches (L1: experimental, many streams)


class ViterbiError(RuntimeError):
    pass


GATHER_NONE, GATHER_NCCL, GATHER_COPY, GATHER_DIRECT = 0, 1, 2, 3
GATHER_MODES = {"none": GATHER_NONE, "nccl": GATHER_NCCL, "copy": GATHER_COPY, "direct": GATHER_DIRECT}
COMM_ID_BYTES = 128


class JobConfig(C.Structure):
    """vit_job_config (include/vit_b200.h)"""
    _fields_ = [("options", C.c_int), ("n_bits", C.c_size_t), ("nstreams", C.c_uint), ("wave", C.c_uint), ("batch", C.c_uint),
                ("seed", C.c_uint), ("source", C.c_int), ("amp", C.c_int), ("sigma", C.c_double), ("gather", C.c_int), ("root", C.c_int)]


class JobResult(C.Structure):
    """vit_job_result (include/vit_b200.h)"""
    _fields_ = [("decode_ms", C.c_double), ("job_ms", C.c_double), ("synth_ms", C.c_double), ("decoded_bits", C.c_ulonglong),
                ("bit_errors", C.c_ulonglong), ("max_stream_errors", C.c_ulonglong), ("launches", C.c_ulonglong), ("streams", C.c_uint)]


_lib = None


def lib():
    """The C-ABI library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ViterbiError("%s is not built: run `make -C %s/csrc` (or __graft_entry__.build())" % (LIB_PATH, _HERE))
        L = C.CDLL(LIB_PATH)
        sz, vp = C.c_size_t, C.c_void_p
        L.vit_create.restype, L.vit_create.argtypes = C.c_int, [C.POINTER(vp), C.c_int, C.c_int, sz]
        L.vit_destroy.restype, L.vit_destroy.argtypes = None, [vp]
        L.vit_run.restype, L.vit_run.argtypes = C.c_int, [vp, vp, vp, sz, C.POINTER(C.c_float)]
        L.vit_run_device.restype = C.c_int
        L.vit_run_device.argtypes = [vp, vp, vp, sz, vp, C.POINTER(C.c_float)]
        L.vit_run_device_batch.restype = C.c_int
        L.vit_run_device_batch.argtypes = [vp, vp, vp, sz, sz, sz, sz, vp, C.POINTER(C.c_float)]
        for n in ("vit_input_size", "vit_message_len", "vit_output_size"):
            f = getattr(L, n)
            f.restype, f.argtypes = sz, [C.c_int, sz]
        L.vit_code_parameters.restype, L.vit_code_parameters.argtypes = None, [C.POINTER(C.c_int)] * 3
        L.vit_options_valid.restype, L.vit_options_valid.argtypes = C.c_int, [C.c_int]
        L.vit_options_valid_ref.restype, L.vit_options_valid_ref.argtypes = C.c_int, [C.c_int]
        L.vit_kernel_info.restype = C.c_int
        L.vit_kernel_info.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 4
        L.vit_launch_count.restype, L.vit_launch_count.argtypes = C.c_ulonglong, [vp]
        L.vit_last_launch_staged_output.restype, L.vit_last_launch_staged_output.argtypes = C.c_int, [vp]
        L.vit_set_segments.restype, L.vit_set_segments.argtypes = C.c_int, [vp, C.c_uint]
        L.vit_last_error.restype, L.vit_last_error.argtypes = C.c_char_p, []
        L.vit_count_errors_device.restype = C.c_int
        L.vit_count_errors_device.argtypes = [C.c_int, vp, vp, sz, C.POINTER(C.c_ulonglong), vp]
        L.vit_dev_alloc.restype, L.vit_dev_alloc.argtypes = C.c_int, [C.POINTER(vp), sz]
        L.vit_dev_free.restype, L.vit_dev_free.argtypes = None, [vp]
        L.vit_dev_sync.restype, L.vit_dev_sync.argtypes = C.c_int, []
        L.vit_dev_count.restype, L.vit_dev_count.argtypes = C.c_int, []
        L.vit_host_alloc.restype, L.vit_host_alloc.argtypes = C.c_int, [C.POINTER(vp), sz]
        L.vit_host_free.restype, L.vit_host_free.argtypes = None, [vp]
        L.vit_synth_device.restype = C.c_int
        L.vit_synth_device.argtypes = [C.c_int, sz, C.c_uint, C.c_int, C.c_double, C.c_int, vp, vp, vp]
        L.vit_set_upload_mode.restype, L.vit_set_upload_mode.argtypes = C.c_int, [vp, C.c_int]
        L.vit_upload_mode_in_effect.restype, L.vit_upload_mode_in_effect.argtypes = C.c_int, [vp]
        L.vit_set_geometry.restype, L.vit_set_geometry.argtypes = C.c_int, [vp, C.c_int]
        L.vit_last_launch_geometry.restype, L.vit_last_launch_geometry.argtypes = C.c_int, [vp]
        L.vit_synth_device_ex.restype = C.c_int
        L.vit_synth_device_ex.argtypes = [C.c_int, sz, C.c_uint, C.c_int, C.c_double, C.c_int, C.c_int, vp, vp, vp]
        L.vit_count_errors_synth_device.restype = C.c_int
        L.vit_count_errors_synth_device.argtypes = [C.c_int, vp, sz, C.c_uint, C.c_int, C.POINTER(C.c_ulonglong), vp]
        L.vit_stream_reset.restype, L.vit_stream_reset.argtypes = C.c_int, [vp]
        L.vit_stream_push.restype, L.vit_stream_push.argtypes = C.c_int, [vp, vp, sz, vp, sz, C.POINTER(sz)]
        L.vit_stream_push_device.restype = C.c_int
        L.vit_stream_push_device.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz), vp]
        L.vit_stream_pending.restype, L.vit_stream_pending.argtypes = sz, [vp]
        L.vit_stream_bits.restype, L.vit_stream_bits.argtypes = C.c_ulonglong, [vp]
        L.vit_depuncture_device.restype = C.c_int
        L.vit_depuncture_device.argtypes = [C.c_int, vp, sz, C.c_uint, C.c_uint, C.c_uint, vp, sz, vp]
        L.vit_dev_set.restype, L.vit_dev_set.argtypes = C.c_int, [C.c_int]
        L.vit_dev_copy_to_host.restype, L.vit_dev_copy_to_host.argtypes = C.c_int, [vp, vp, sz]
        L.vit_dev_copy_from_host.restype, L.vit_dev_copy_from_host.argtypes = C.c_int, [vp, vp, sz]
        L.vit_comm_available.restype, L.vit_comm_available.argtypes = C.c_int, []
        L.vit_comm_nccl_version.restype, L.vit_comm_nccl_version.argtypes = C.c_int, []
        L.vit_comm_get_unique_id.restype, L.vit_comm_get_unique_id.argtypes = C.c_int, [vp]
        L.vit_comm_init_rank.restype, L.vit_comm_init_rank.argtypes = C.c_int, [C.POINTER(vp), C.c_int, C.c_int, vp, C.c_int]
        L.vit_comm_init_all.restype, L.vit_comm_init_all.argtypes = C.c_int, [C.POINTER(vp), C.c_int, C.POINTER(C.c_int)]
        L.vit_comm_destroy.restype, L.vit_comm_destroy.argtypes = None, [vp]
        L.vit_comm_rank.restype, L.vit_comm_rank.argtypes = C.c_int, [vp]
        L.vit_comm_size.restype, L.vit_comm_size.argtypes = C.c_int, [vp]
        L.vit_comm_barrier.restype, L.vit_comm_barrier.argtypes = C.c_int, [vp]
        L.vit_comm_stream_wait.restype, L.vit_comm_stream_wait.argtypes = C.c_int, [vp, vp]
        L.vit_comm_mark.restype, L.vit_comm_mark.argtypes = C.c_int, [vp, C.c_int]
        L.vit_comm_stream_wait_mark.restype, L.vit_comm_stream_wait_mark.argtypes = C.c_int, [vp, C.c_int, vp]
        L.vit_comm_stream.restype, L.vit_comm_stream.argtypes = vp, [vp]
        L.vit_shard_range.restype, L.vit_shard_range.argtypes = None, [sz, C.c_int, C.c_int, C.POINTER(sz), C.POINTER(sz)]
        L.vit_shard_owner.restype, L.vit_shard_owner.argtypes = C.c_int, [sz, C.c_int, sz]
        L.vit_comm_shared_alloc.restype, L.vit_comm_shared_alloc.argtypes = C.c_int, [vp, C.POINTER(vp), sz, C.c_int]
        L.vit_comm_gatherv.restype = C.c_int
        L.vit_comm_gatherv.argtypes = [vp, C.c_int, vp, vp, C.POINTER(sz), C.POINTER(sz), C.c_int, vp]
        L.vit_job_create.restype, L.vit_job_create.argtypes = C.c_int, [C.POINTER(vp), vp, C.c_int, C.POINTER(JobConfig)]
        L.vit_job_run.restype, L.vit_job_run.argtypes = C.c_int, [vp, C.POINTER(JobResult)]
        L.vit_job_gathered.restype, L.vit_job_gathered.argtypes = vp, [vp, C.POINTER(sz)]
        L.vit_job_stream_range.restype, L.vit_job_stream_range.argtypes = C.c_int, [vp, C.POINTER(sz), C.POINTER(sz)]
        L.vit_job_stream_errors.restype, L.vit_job_stream_errors.argtypes = C.c_int, [vp, C.POINTER(C.c_ulonglong), sz]
        L.vit_job_destroy.restype, L.vit_job_destroy.argtypes = None, [vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise ViterbiError("vit error %d: %s" % (rc, lib().vit_last_error().decode()))


def code_parameters():
    """(constLen, polyn1, polyn2) this build of the library decodes (vit_code_parameters; reference viterbi.h:61-63)."""
    v = [C.c_int(0) for _ in range(3)]
    lib().vit_code_parameters(*[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def options_valid(options):
    return bool(lib().vit_options_valid(options))


def options_valid_ref(options):
    return bool(lib().vit_options_valid_ref(options))


def parse_options(input="h", metric="b32", output="b32", comp="reg"):
    """The ./main flag values (reference src/main.cpp:211-254) -> option bitfield."""
    i = {"HARD": HARD, "h": HARD, "SOFT4": SOFT4, "s4": SOFT4, "SOFT8": SOFT8, "s8": SOFT8,
         "SOFT16": SOFT16, "s16": SOFT16, "FP32": FP32, "f": FP32}[input]
    m = {"b16": M_B16, "b32": M_B32, "f16": M_FP16}[metric]
    o = {"b16": O_B16, "b32": O_B32}[output]
    c = {"REG": REG, "reg": REG, "DPX": DPX, "dpx": DPX, "DPXT": DPX_TIES, "dpxt": DPX_TIES}[comp]
    return i | m | o | c


SOURCE_HASH, SOURCE_PRBS31 = 0, 1


def synth_device(input_type, n_bits, packed_ptr, bits_ptr=None, seed=1, amp=0, sigma=0.0, zero=False, stream=0,
                 source=SOURCE_HASH):
    """Device-side synthetic received stream (vit_synth_device_ex): the GPU twin of the reference harness's host
    source/encoder/noise/packer chain.  packed_ptr must hold whole 32-bit packs.  source: counter-hash message bits or
    PRBS-31 started from state `seed`."""
    _check(lib().vit_synth_device_ex(int(input_type), int(n_bits), int(seed), int(amp), float(sigma), int(bool(zero)),
                                     int(source), packed_ptr, bits_ptr, stream))


def count_errors_synth_device(options, out_ptr, message_len, seed=1, source=SOURCE_HASH, stream=0):
    """Bit errors against the synthetic source's own message bits, regenerated on the device (no bits buffer)."""
    n = C.c_ulonglong(0)
    _check(lib().vit_count_errors_synth_device(int(options), out_ptr, int(message_len), int(seed), int(source), C.byref(n), stream))
    return int(n.value)


def count_errors_device(options, out_ptr, bits_ptr, message_len, stream=0):
    """Bit errors counted on the device (vit_count_errors_device): out bit j vs message bit j+26."""
    n = C.c_ulonglong(0)
    _check(lib().vit_count_errors_device(int(options), out_ptr, bits_ptr, int(message_len), C.byref(n), stream))
    return int(n.value)


class ViterbiCUDA:
    """Mirror of reference `ViterbiCUDA<options>` (viterbi.h:91-152)."""

    def __init__(self, options=0, inputNum=0, device=0):
        self.options = int(options)
        self._h = C.c_void_p()
        _check(lib().vit_create(C.byref(self._h), self.options, int(device), int(inputNum)))
        self.device = int(device)
        self.bitsPerPack = 16 if (self.options & DECODE_MASK) == O_B16 else 32
        self.decPack_t = np.uint16 if self.bitsPerPack == 16 else np.uint32
        self.encPack_t = np.float32 if (self.options & CHANNEL_MASK) == FP32 else np.int32

    def close(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.vit_destroy(h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass

    # reference viterbi.cu:63-92
    def getInputSize(self, inputNum):
        return lib().vit_input_size(self.options, inputNum)

    def getMessageLen(self, inputNum):
        return lib().vit_message_len(self.options, inputNum)

    def getOutputSize(self, inputNum):
        return lib().vit_output_size(self.options, inputNum)

    def run(self, input_h, inputNum, output_h=None, want_kernel_time=False):
        """reference ViterbiCUDA::run (viterbi.cu:210-238): host numpy buffers in and out.
        Returns output_h, or (output_h, kernel_ms) when want_kernel_time."""
        input_h = np.ascontiguousarray(input_h)
        if input_h.nbytes < self.getInputSize(inputNum):
            raise ViterbiError("input buffer holds %d bytes, %d needed" % (input_h.nbytes, self.getInputSize(inputNum)))
        nwords = self.getOutputSize(inputNum) // np.dtype(self.decPack_t).itemsize
        if output_h is None:
            output_h = np.empty(nwords, self.decPack_t)
        elif (not isinstance(output_h, np.ndarray) or output_h.nbytes < self.getOutputSize(inputNum)
              or not output_h.flags["C_CONTIGUOUS"] or not output_h.flags["WRITEABLE"]):
            raise ViterbiError("output_h must be a writable C-contiguous numpy array of at least %d bytes"
                               % self.getOutputSize(inputNum))
        ms = C.c_float(0)
        _check(lib().vit_run(self._h, input_h.ctypes.data, output_h.ctypes.data, inputNum,
                             C.byref(ms) if want_kernel_time else None))
        return (output_h, ms.value) if want_kernel_time else output_h

    def run_device(self, in_ptr, out_ptr, inputNum, stream=0, want_kernel_time=False,
                   nstreams=1, in_stride=0, out_stride=0):
        """Device-resident decode: raw device pointers (e.g. torch tensor .data_ptr())."""
        ms = C.c_float(0)
        _check(lib().vit_run_device_batch(self._h, in_ptr, out_ptr, inputNum, nstreams, in_stride, out_stride,
                                          stream, C.byref(ms) if want_kernel_time else None))
        return ms.value if want_kernel_time else None

    # chunked decode of an endless stream (vit_stream_*): the reference has no equivalent (viterbi.cu:168-169)
    def stream_reset(self):
        _check(lib().vit_stream_reset(self._h))

    def stream_push(self, input_h, inputNum):
        """Append inputNum coded symbols (whole 32-bit channel packs); returns the decoded packs this chunk completes."""
        input_h = np.ascontiguousarray(input_h)
        if input_h.nbytes < self.getInputSize(inputNum):
            raise ViterbiError("input buffer holds %d bytes, %d needed" % (input_h.nbytes, self.getInputSize(inputNum)))
        cap = self.getOutputSize(self.stream_pending() + inputNum)
        out = np.empty(cap // np.dtype(self.decPack_t).itemsize, self.decPack_t)
        n = C.c_size_t(0)
        _check(lib().vit_stream_push(self._h, input_h.ctypes.data, int(inputNum), out.ctypes.data, cap, C.byref(n)))
        return out[: n.value // np.dtype(self.decPack_t).itemsize]

    def stream_pending(self):
        return int(lib().vit_stream_pending(self._h))

    def stream_bits(self):
        return int(lib().vit_stream_bits(self._h))

    def set_upload_mode(self, mode):
        """UPLOAD_AUTO / _SEQUENTIAL / _CHUNKED / _GATED: how run() moves host buffers (vit_set_upload_mode)."""
        _check(lib().vit_set_upload_mode(self._h, int(mode)))

    def set_geometry(self, geometry):
        """GEOMETRY_L8 (default) / GEOMETRY_L1 (experimental: one lane per segment, for launches of a dozen streams and more)"""
        _check(lib().vit_set_geometry(self._h, int(geometry)))

    def last_launch_geometry(self):
        return int(lib().vit_last_launch_geometry(self._h))

    def upload_mode_in_effect(self):
        return int(lib().vit_upload_mode_in_effect(self._h))

    def kernel_info(self):
        v = [C.c_int(0) for _ in range(4)]
        _check(lib().vit_kernel_info(self.options, *[C.byref(x) for x in v]))
        return {"regs": v[0].value, "smem_bytes": v[1].value, "block_threads": v[2].value, "segs_per_block": v[3].value}

    def last_launch_staged_output(self):
        return bool(lib().vit_last_launch_staged_output(self._h))

    def launch_count(self):
        return int(lib().vit_launch_count(self._h))

    def set_segments(self, segments):
        _check(lib().vit_set_segments(self._h, segments))


# ---------------------------------------------------------------------------------------------------------------------
# multi-GPU: stream sharding + gather of the packed output bits (C ABI vit_comm_* / vit_job_*, csrc/vit_mg.cu)

# standard puncturing patterns of the K=7 (171,133) mother code (DVB-S / 802.11): (period, keep0, keep1)
PUNCTURE = {"1/2": (1, 0b1, 0b1), "2/3": (2, 0b01, 0b11), "3/4": (3, 0b101, 0b011), "5/6": (5, 0b10101, 0b01011), "7/8": (7, 0b1010001, 0b0101111)}


def depuncture_device(input_type, in_ptr, n_in_syms, rate, out_ptr, n_out_stages, stream=0):
    """Expand a punctured soft-symbol stream to rate 1/2 with erasures (vit_depuncture_device).  rate: a key of PUNCTURE or
    a (period, keep0, keep1) tuple."""
    period, k0, k1 = PUNCTURE[rate] if isinstance(rate, str) else rate
    _check(lib().vit_depuncture_device(int(input_type), in_ptr, int(n_in_syms), int(period), int(k0), int(k1), out_ptr,
                                       int(n_out_stages), stream))


def dev_to_host(src_ptr, nbytes):
    """nbytes of device memory at src_ptr as a numpy uint8 array (synchronous copy)."""
    out = np.empty(int(nbytes), np.uint8)
    _check(lib().vit_dev_copy_to_host(out.ctypes.data, src_ptr, int(nbytes)))
    return out


def shard_range(nstreams, nranks, rank):
    """(first, count): the contiguous block of streams rank owns (vit_shard_range)."""
    a, b = C.c_size_t(0), C.c_size_t(0)
    lib().vit_shard_range(int(nstreams), int(nranks), int(rank), C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def shard_owner(nstreams, nranks, stream):
    return int(lib().vit_shard_owner(int(nstreams), int(nranks), int(stream)))


def comm_unique_id():
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(lib().vit_comm_get_unique_id(buf))
    return buf.raw


class Comm:
    """One rank of an N-GPU communicator (NCCL underneath, loaded at run time)."""

    def __init__(self, nranks, rank, unique_id, device):
        self._c = C.c_void_p()
        idb = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        _check(lib().vit_comm_init_rank(C.byref(self._c), int(nranks), int(rank), idb, int(device)))
        self.rank, self.nranks, self.device = int(rank), int(nranks), int(device)

    def barrier(self):
        _check(lib().vit_comm_barrier(self._c))

    def stream_wait(self, stream):
        _check(lib().vit_comm_stream_wait(self._c, stream))

    def mark(self, k):
        _check(lib().vit_comm_mark(self._c, int(k)))

    def stream_wait_mark(self, k, stream):
        _check(lib().vit_comm_stream_wait_mark(self._c, int(k), stream))

    def shared_alloc(self, nbytes, root=0):
        p = C.c_void_p()
        _check(lib().vit_comm_shared_alloc(self._c, C.byref(p), int(nbytes), int(root)))
        return p.value

    def size_array(self, values):
        """a per-rank size_t array for gatherv (build once, reuse every step)"""
        return (C.c_size_t * self.nranks)(*[int(x) for x in values])

    def gatherv(self, mode, send_ptr, recv_base, offsets, sizes, root=0, producer_stream=0):
        off = offsets if isinstance(offsets, C.Array) else self.size_array(offsets)
        siz = sizes if isinstance(sizes, C.Array) else self.size_array(sizes)
        _check(lib().vit_comm_gatherv(self._c, int(mode), send_ptr, recv_base, off, siz, int(root), producer_stream))

    def close(self):
        if getattr(self, "_c", None) and _lib is not None:
            _lib.vit_comm_destroy(self._c)
            self._c = C.c_void_p()


class StreamJob:
    """The sharded stream job (vit_job_*): nstreams independent streams generated, decoded and gathered on the GPUs."""

    def __init__(self, comm, device, **cfg):
        self.cfg = JobConfig(**cfg)
        self._j = C.c_void_p()
        self._comm = comm
        _check(lib().vit_job_create(C.byref(self._j), comm._c if comm is not None else None, int(device), C.byref(self.cfg)))

    def run(self):
        r = JobResult()
        _check(lib().vit_job_run(self._j, C.byref(r)))
        return {k: getattr(r, k) for k, _ in JobResult._fields_}

    def gathered(self):
        st = C.c_size_t(0)
        p = lib().vit_job_gathered(self._j, C.byref(st))
        return p, int(st.value)

    def stream_range(self):
        a, b = C.c_size_t(0), C.c_size_t(0)
        _check(lib().vit_job_stream_range(self._j, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def stream_errors(self):
        first, count = self.stream_range()
        arr = (C.c_ulonglong * max(count, 1))()
        _check(lib().vit_job_stream_errors(self._j, arr, count))
        return [int(arr[i]) for i in range(count)]

    def close(self):
        if getattr(self, "_j", None) and _lib is not None:
            _lib.vit_job_destroy(self._j)
            self._j = C.c_void_p()
