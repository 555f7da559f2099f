"""CPU, world_size 2 over gloo: the host-side sharding / gather logic of metrabs_b200.parallel.  The device stages
are replaced by the oracle port (test infrastructure) so the test checks exactly what the N>1 path adds: contiguous
ragged chunks, rank-ordered gather of [coords2d|coords3d_rel], and full-batch reconstruction on every rank, which
must reproduce the UNSHARDED result (batch-global RMS, ptu3d.py:71-74) exactly."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_no, n_total, out_dir):
    sys.path.insert(0, ROOT)
    from metrabs_b200 import parallel
    from oracle import port
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port_no)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    pcfg = port.PathConfig(proc_side=64)
    spec = port.effnet_spec('efficientnetv2-tiny')
    sd = port.make_effnet_state_dict(spec, pcfg, 8, seed=0)
    crops, k = port.synthetic_inputs(n_total, 64, seed=3)
    s, e = parallel.shard_range(n_total, world, rank)
    with torch.inference_mode():
        feats = port.effnet_features(sd, spec, crops[s:e])
        c2d, c3d = port.heads(sd, feats, pcfg, 8)

        def all_gather(t):
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            return torch.stack(outs)

        packed = parallel.gather_decoded(parallel.pack_decoded(c2d, c3d), n_total, world, all_gather)
        g2d, g3d = parallel.unpack_decoded(packed)
        out = port.reconstruct_absolute(g2d, g3d, k, pcfg)
        ref = port.metrabs_forward(sd, spec, pcfg, 8, crops, k)
    torch.save(dict(out=out, ref=ref), os.path.join(out_dir, f'r{rank}.pt'))
    dist.destroy_process_group()


class _OracleEngine:
    """Engine-shaped adapter over the oracle port + gloo, so ShardedMetrabs.forward itself runs on the CPU."""

    def __init__(self, sd, spec, pcfg, world):
        self.sd, self.spec, self.pcfg, self.world, self.n_joints = sd, spec, pcfg, world, 8

    def backbone(self, crops):
        from oracle import port
        return port.effnet_features(self.sd, self.spec, crops)

    def head_decode(self, feats):
        from oracle import port
        return port.heads(self.sd, feats, self.pcfg, 8)

    def allgather(self, t):
        outs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(outs, t)
        return torch.stack(outs)

    def reconstruct_absolute(self, c2d, c3d, k):
        from oracle import port
        return port.reconstruct_absolute(c2d, c3d, k, self.pcfg)


def _worker_sharded_forward(rank, world, port_no, n_total, out_dir):
    sys.path.insert(0, ROOT)
    from metrabs_b200 import parallel
    from oracle import port
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port_no)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    pcfg = port.PathConfig(proc_side=64)
    spec = port.effnet_spec('efficientnetv2-tiny')
    sd = port.make_effnet_state_dict(spec, pcfg, 8, seed=0)
    crops, k = port.synthetic_inputs(n_total, 64, seed=3)
    with torch.inference_mode():
        out = parallel.ShardedMetrabs(None, rank, world, engine=_OracleEngine(sd, spec, pcfg, world)).forward(crops, k)
        ref = port.metrabs_forward(sd, spec, pcfg, 8, crops, k)
    torch.save(dict(out=out, ref=ref), os.path.join(out_dir, f'r{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [1, 3])
def test_sharded_forward_with_empty_shards(tmp_path, n_total):
    """Fewer crops than ranks (a normal load for the multiperson caller): ranks with an empty shard skip the device
    stages but still enter the all-gather; nobody hangs, every rank returns the unsharded result."""
    world = 2 if n_total == 1 else 4
    port_no = 31500 + (os.getpid() % 2000) + n_total
    mp.spawn(_worker_sharded_forward, args=(world, port_no, n_total, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f'r{r}.pt') for r in range(world)]
    for o in outs:
        assert o['out'].shape == (n_total, 8, 3)
        # (torch-cpu picks batch-size-dependent conv algorithms: a 1-crop shard and the 3-crop reference differ in summation order)
        assert (o['out'] - o['ref']).abs().max() / o['ref'].abs().max() < 2e-4
        assert (o['out'] - outs[0]['out']).abs().max() / o['ref'].abs().max() < 1e-6  # every rank solved the same full batch


@pytest.mark.parametrize('n_total', [5, 8])
def test_sharded_equals_unsharded(tmp_path, n_total):
    world = 2
    port_no = 29500 + (os.getpid() % 2000) + n_total
    mp.spawn(_worker, args=(world, port_no, n_total, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f'r{r}.pt') for r in range(world)]
    for o in outs:
        assert o['out'].shape == (n_total, 8, 3)
        # per-crop stages are batch-independent up to conv summation order; the reconstruction sees the full batch
        assert (o['out'] - o['ref']).abs().max() / o['ref'].abs().max() < 1e-5
    assert torch.equal(outs[0]['out'], outs[1]['out'])


def test_shard_ranges_cover_batch():
    from metrabs_b200 import parallel
    for n in (1, 7, 8, 256, 257):
        for w in (1, 2, 4, 8):
            ranges = [parallel.shard_range(n, w, r) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in ranges) - min(e - s for s, e in ranges) <= 1
