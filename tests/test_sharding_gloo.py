"""CPU test of the N>1 path: world_size-2 gloo run of the stream sharding + output gather.  The per-rank
decode is stood in for by the golden model (test infrastructure); the partition and gather code is the
product's (gpu-accelerated-viterbi-decoder_b200/sharding.py), the same code bench.py runs over NCCL."""
import os
import subprocess
import sys
import textwrap

import pytest

from vit_testlib import ROOT

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    from vit_testlib import load_pkg
    from oracle import oracle as O
    import importlib.util
    spec = importlib.util.spec_from_file_location("vit_sharding", os.path.join(%(root)r, "gpu-accelerated-viterbi-decoder_b200", "sharding.py"))
    S = importlib.util.module_from_spec(spec); spec.loader.exec_module(S)
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_streams, n_bits, opt = 5, 64 + 16 * 900, 0x112
    mine = S.streams_of_rank(n_streams, world, rank)
    outs = []
    for s in mine:
        bits, packed, N = O.make_channel_det(n_bits, O.SOFT8, seed=100 + s, sigma=0.9)
        outs.append(torch.from_numpy(O.decode(opt, packed, N).astype(np.int32)))
    local = torch.stack(outs) if outs else torch.zeros((0, O.output_size(opt, 2 * n_bits) // 2), dtype=torch.int32)
    full = S.gather_packed_outputs(dist, local, n_streams, world, rank)
    assert full.shape[0] == n_streams
    for s in range(n_streams):
        bits, packed, N = O.make_channel_det(n_bits, O.SOFT8, seed=100 + s, sigma=0.9)
        assert np.array_equal(full[s].numpy().astype(np.uint16), O.decode(opt, packed, N)), (rank, s)
        assert S.owner_of_stream(n_streams, world, s) == [k for k in range(world) if s in S.streams_of_rank(n_streams, world, k)][0]
    dist.barrier()
    if rank == 0:
        print("SHARDING_OK")
    dist.destroy_process_group()
''')


def test_partition_is_a_partition():
    import importlib.util
    spec = importlib.util.spec_from_file_location("vit_sharding", os.path.join(ROOT, "gpu-accelerated-viterbi-decoder_b200", "sharding.py"))
    S = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(S)
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


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "SHARDING_OK" in out.stdout
