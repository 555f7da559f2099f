"""Run under torchrun (one rank per GPU): the sharded crop-model forward (metrabs_b200.parallel.ShardedMetrabs: contiguous
chunks, ONE NCCL all-gather of [coords2d|coords3d_rel] through mtb_allgather_joints, full-batch reconstruction on every
rank) must equal the single-GPU forward of the whole batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse

import torch
import torch.distributed as dist

import bench
from metrabs_b200.parallel import ShardedMetrabs

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
for precision in ('fp32', 'bf16'):
    args = argparse.Namespace(size='s', side=256, joints=24, precision=precision)
    model = bench.build_model(args, dev)
    eng = model.engine(dev)

    def bcast(raw):
        t = torch.tensor(list(raw) if raw is not None else [0] * 128, dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    eng.comm_init(rank, world, bcast)
    n = 13  # ragged over the ranks
    crops, k = bench.synthetic(n, 256, 5)
    crops, k = crops.to(dev), k.to(dev)
    full = eng.forward(crops, k)
    sharded = ShardedMetrabs(model, rank, world).forward(crops, k)
    torch.cuda.synchronize()
    err = float((full - sharded).abs().max() / full.abs().max())
    same = torch.equal(full, sharded)
    print(f'rank {rank}/{world} {precision}: sharded vs unsharded max rel diff {err:.2e} bit-equal={same}', flush=True)
    assert err < (1e-6 if precision == 'fp32' else 2e-2), err
dist.destroy_process_group()
