"""Multi-GPU stream sharding: independent codeword streams are dealt to ranks in contiguous blocks (one process or
thread per GPU), each rank decodes its own streams with its own decoder, and ONLY the packed output bits are gathered.
A stream is never split across ranks -- the reference has no multi-GPU path at all (cudaSetDevice(0), viterbi.cu:134).

The partition is the library's own (C ABI vit_shard_range / vit_shard_owner in csrc/vit_mg.cu, used by the stream job
vit_job_run and by bench.py); these helpers only re-shape it for Python callers and for the gloo-backed CPU test of the
gather protocol (every rank derives every rank's offset and size from the partition alone, as vit_job_run does).
"""
import importlib
import sys


def _pkg():
    return sys.modules[__name__.rsplit(".", 1)[0]] if "." in __name__ else importlib.import_module("gpu_accelerated_viterbi_decoder_b200")


def streams_of_rank(n_streams, world, rank):
    first, count = _pkg().shard_range(n_streams, world, rank)
    return list(range(first, first + count))


def owner_of_stream(n_streams, world, s):
    return _pkg().shard_owner(n_streams, world, s)


def wave_blocks(n_streams, world, batch, wave, out_stride, round_idx, wave0):
    """(offsets, sizes) in bytes, one entry per rank, of the blocks gathered after the wave starting at stream `wave0`
    of batch round `round_idx` -- what vit_job_run passes to vit_comm_gatherv."""
    offsets, sizes = [], []
    b0 = round_idx * batch
    for p in range(world):
        pf, pc = _pkg().shard_range(n_streams, world, p)
        pnb = min(batch, pc - b0) if b0 < pc else 0
        pnw = min(wave, pnb - wave0) if wave0 < pnb else 0
        offsets.append((pf + b0 + wave0) * out_stride)
        sizes.append(pnw * out_stride)
    return offsets, sizes
