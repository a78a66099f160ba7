"""CPU test of the N>1 path: world_size-2 gloo run of the stream sharding + gather-to-root protocol.  The per-rank
decode is stood in for by the golden model (test infrastructure) and gloo send/recv stands in for NCCL send/recv; the
partition (C ABI vit_shard_range / vit_shard_owner, csrc/vit_mg.cu) and the per-wave offsets/sizes every rank derives
from it (sharding.wave_blocks == what vit_job_run passes to vit_comm_gatherv) are the product's."""
import os
import subprocess
import sys
import textwrap

import pytest

from vit_testlib import ROOT, load_pkg

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    from vit_testlib import load_pkg
    from oracle import oracle as O
    V = load_pkg()
    import importlib.util
    spec = importlib.util.spec_from_file_location("vit_sharding", os.path.join(%(root)r, "gpu-accelerated-viterbi-decoder_b200", "sharding.py"))
    S = importlib.util.module_from_spec(spec); spec.loader.exec_module(S)
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_streams, n_bits, opt, batch, wave, root = 7, 64 + 16 * 900, 0x112, 2, 2, 0
    out_bytes = O.output_size(opt, 2 * n_bits)
    out_stride = (out_bytes + 255) // 256 * 256
    mine = S.streams_of_rank(n_streams, world, rank)
    gathered = torch.zeros(n_streams * out_stride, dtype=torch.uint8) if rank == root else None
    max_count = len(S.streams_of_rank(n_streams, world, 0))
    for rnd in range((max_count + batch - 1) // batch):
        for w0 in range(0, batch, wave):
            offsets, sizes = S.wave_blocks(n_streams, world, batch, wave, out_stride, rnd, w0)
            # this rank's block of the wave: decode its streams (golden model stands in for the GPU)
            block = torch.zeros(sizes[rank], dtype=torch.uint8)
            for k in range(sizes[rank] // out_stride):
                s = offsets[rank] // out_stride + k
                assert s in mine
                bits, packed, N = O.make_channel_det(n_bits, O.SOFT8, seed=100 + s, sigma=0.9)
                block[k * out_stride:k * out_stride + out_bytes] = torch.from_numpy(O.decode(opt, packed, N).view(np.uint8).copy())
            if rank == root:
                gathered[offsets[root]:offsets[root] + sizes[root]] = block
                for p in range(world):
                    if p != root and sizes[p]:
                        buf = torch.zeros(sizes[p], dtype=torch.uint8)
                        dist.recv(buf, src=p)
                        gathered[offsets[p]:offsets[p] + sizes[p]] = buf
            elif sizes[rank]:
                dist.send(block, dst=root)
    dist.barrier()
    if rank == root:
        for s in range(n_streams):
            bits, packed, N = O.make_channel_det(n_bits, O.SOFT8, seed=100 + s, sigma=0.9)
            got = gathered[s * out_stride:s * out_stride + out_bytes].numpy().view(np.uint16)
            assert np.array_equal(got, O.decode(opt, packed, N)), s
            assert S.owner_of_stream(n_streams, world, s) == [k for k in range(world) if s in S.streams_of_rank(n_streams, world, k)][0]
        print("SHARDING_OK")
    dist.destroy_process_group()
''')


def _sharding():
    import importlib.util
    load_pkg()
    spec = importlib.util.spec_from_file_location("vit_sharding", os.path.join(ROOT, "gpu-accelerated-viterbi-decoder_b200", "sharding.py"))
    S = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(S)
    return S


def test_partition_is_a_partition():
    S = _sharding()
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                seen += S.streams_of_rank(n, world, r)
            assert seen == list(range(n))
            sizes = [len(S.streams_of_rank(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
            for s in range(n):
                assert s in S.streams_of_rank(n, world, S.owner_of_stream(n, world, s))


def test_wave_blocks_tile_the_gathered_buffer():
    """Over all rounds and waves the per-rank blocks cover every stream exactly once, in stream-major order."""
    S = _sharding()
    for n, world, batch, wave in ((1024, 8, 32, 16), (7, 2, 2, 2), (5, 4, 4, 2), (9, 3, 2, 1)):
        stride = 256
        covered = []
        max_count = len(S.streams_of_rank(n, world, 0))
        for rnd in range((max_count + batch - 1) // batch):
            for w0 in range(0, batch, wave):
                offsets, sizes = S.wave_blocks(n, world, batch, wave, stride, rnd, w0)
                for p in range(world):
                    for k in range(sizes[p] // stride):
                        s = offsets[p] // stride + k
                        assert S.owner_of_stream(n, world, s) == p
                        covered.append(s)
        assert sorted(covered) == list(range(n)), (n, world, batch, wave)


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "SHARDING_OK" in out.stdout
