"""GPU, >= 2 devices (skipped otherwise): the data-parallel path over NCCL - mtb_forward_sharded (local backbone + head
decode, ONE ncclAllGather of [coords2d|coords3d_rel], full-batch reconstruction) must reproduce the UNSHARDED forward of
the concatenated batch on every rank (batch-global RMS, ptu3d.py:71-74), including ragged and empty shards."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_no, precision, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port_no)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from metrabs_b200 import parallel
    from oracle import port
    from tests import helpers
    pcfg = port.PathConfig(proc_side=64)
    sd = port.make_effnet_state_dict(port.effnet_spec('efficientnetv2-tiny'), pcfg, 8, seed=0)
    m = helpers.device_model('efficientnetv2-tiny', pcfg, 8, sd, precision=precision).to(dev)
    eng = m.engine(dev)

    def bcast(raw):
        t = torch.tensor(list(raw) if raw is not None else [0] * 128, dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())
    eng.comm_init(rank, world, bcast)
    sh = parallel.ShardedMetrabs(m, rank, world)
    res = {}
    for n_total in (8, 5, 1):  # equal shards (library path), ragged, fewer crops than ranks
        crops, k = port.synthetic_inputs(n_total, 64, seed=3)
        crops, k = crops.to(dev), k.to(dev)
        out = sh.forward(crops, k)
        ref = eng.forward(crops, k)
        torch.cuda.synchronize()
        res[n_total] = (out.cpu(), ref.cpu())
    torch.save(res, os.path.join(out_dir, f'r{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_sharded_equals_unsharded_nccl(tmp_path, precision):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 CUDA devices')
    world = 2
    port_no = 33500 + (os.getpid() % 2000) + (0 if precision == 'fp32' else 1)
    mp.spawn(_worker, args=(world, port_no, precision, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f'r{r}.pt') for r in range(world)]
    tol = 1e-6 if precision == 'fp32' else 2e-2
    for n_total in (8, 5, 1):
        for r in range(world):
            out, ref = outs[r][n_total]
            assert out.shape == (n_total, 8, 3)
            err = float((out - ref).abs().max() / ref.abs().max())
            assert err <= tol, (n_total, r, err)
        assert torch.equal(outs[0][n_total][0], outs[1][n_total][0])  # every rank holds the same full result
