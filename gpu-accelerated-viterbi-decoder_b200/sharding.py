"""Multi-GPU stream sharding: independent codeword streams are dealt to ranks (one process per GPU),
each rank decodes its own streams with its own persistent decoder, and ONLY the packed output bits are
gathered (NCCL over NVLink on GPUs; the same code runs on gloo for CPU tests).  A stream is never split
across ranks -- the reference has no multi-GPU path at all (cudaSetDevice(0), viterbi.cu:134).
"""


def streams_of_rank(n_streams, world, rank):
    """Contiguous block partition: rank r owns streams [lo, hi).  Blocks differ by at most one stream."""
    q, r = divmod(n_streams, world)
    lo = q * rank + min(rank, r)
    return list(range(lo, lo + q + (1 if rank < r else 0)))


def owner_of_stream(n_streams, world, s):
    q, r = divmod(n_streams, world)
    edge = (q + 1) * r
    return s // (q + 1) if s < edge else r + (s - edge) // max(q, 1)


def gather_packed_outputs(dist, local_outputs, n_streams, world, rank):
    """local_outputs: tensor [n_local, words] of packed decoded bits for this rank's streams, in stream
    order.  Returns the [n_streams, words] tensor on every rank (all_gather of equal-sized, padded blocks)."""
    import torch
    q, r = divmod(n_streams, world)
    cap = q + (1 if r else 0)
    words = local_outputs.shape[1]
    block = torch.zeros((cap, words), dtype=local_outputs.dtype, device=local_outputs.device)
    block[: local_outputs.shape[0]] = local_outputs
    blocks = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(blocks, block)
    parts = [blocks[k][: len(streams_of_rank(n_streams, world, k))] for k in range(world)]
    return torch.cat(parts, 0)
