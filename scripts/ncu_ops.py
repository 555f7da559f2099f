"""Runs named backbone ops in isolation (mtb_debug_run_op) so that ncu can capture exactly those launches.
Usage: ncu --set full -k regex:<kernel> -c N python scripts/ncu_ops.py --precision tf32x3 --batch 128 --ops a,b,c [--reps 1]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--size', default='l')
ap.add_argument('--side', type=int, default=256)
ap.add_argument('--batch', type=int, default=128)
ap.add_argument('--joints', type=int, default=24)
ap.add_argument('--precision', default='bf16')
ap.add_argument('--ops', required=True)
ap.add_argument('--reps', type=int, default=1)
ap.add_argument('--fused', action='store_true', help='run the fused FusedMBConv block (fmb_kernel) that starts at each op')
args = ap.parse_args()
dev = torch.device('cuda', 0)
model = bench.build_model(args, dev)
eng = model.engine(dev)
names = eng.op_names()
g = torch.Generator().manual_seed(0)
for nm in args.ops.split(','):
    i = names.index(nm)
    io = eng.op_io(i)
    x = torch.randn((args.batch,) + io['in_shape'], generator=g).to(dev)
    res = torch.randn((args.batch,) + io['out_shape'], generator=g).to(dev) if io['residual'] else None
    sc = torch.rand(args.batch, io['in_shape'][2], generator=g).to(dev) if io['scale'] else None
    for _ in range(args.reps):
        if args.fused:
            eng.debug_run_fused_block(i, x)
        else:
            eng.debug_run_op(i, x, res, sc)
    torch.cuda.synchronize()
    print('ran', nm, io)
