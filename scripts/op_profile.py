"""Per-op device-time table of one forward (CUDA events around every launch): achieved TFLOP/s and GB/s per op.
Usage: python scripts/op_profile.py [--size l] [--side 256] [--batch 128] [--precision bf16] [--top 40]"""
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
ap.add_argument('--top', type=int, default=45)
args = ap.parse_args()
dev = torch.device('cuda', 0)
model = bench.build_model(args, dev)
eng = model.engine(dev)
crops, k = bench.synthetic(args.batch, args.side, 0)
crops, k = crops.to(dev), k.to(dev)
for _ in range(3):
    eng.forward(crops, k)
torch.cuda.synchronize()
eng.profile_begin(None)
eng.forward(crops, k)
torch.cuda.synchronize()
classes = eng.profile_end()
rows = eng.profile_op_times()
tot = sum(r[2] for r in rows)
print(f'total backbone op time {tot:.3f} ms for {args.batch} crops; classes:',
      {n: round(v["ms"], 3) for n, v in classes.items()})
agg = {}
for name, cls, ms, fl, by_act, by_w in rows:
    by = by_act + by_w / args.batch  # bytes per crop of a launch on `batch` crops: activations + the weights' share
    key = (cls, round(fl), round(by))
    a = agg.setdefault(key, [name, 0, 0.0, fl, by])
    a[1] += 1
    a[2] += ms
pk = bench.peaks()
# floor = the larger of the op's algorithmic FLOPs at the measured tensor peak and its algorithmic bytes at the measured
# HBM peak; x-floor = measured time / floor (1.0 = on the roofline), floor-ms = what the shape would cost on it
print(f'{"first op of shape":44s} {"class":26s} {"n":>3s} {"ms":>8s} {"%":>5s} {"TFLOP/s":>8s} {"GB/s":>7s} {"MFLOP/crop":>10s} '
      f'{"KB/crop":>8s} {"floor-ms":>8s} {"x-floor":>7s}')
floor_tot = 0.0
for (cls, _, _), (name, n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    floor_tot += n * max(fl * args.batch / (pk['tflops'] * 1e12), by * args.batch / (pk['hbm_gbs'] * 1e9)) * 1e3
for (cls, _, _), (name, n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:args.top]:
    per = ms / n / 1e3
    floor = n * max(fl * args.batch / (pk['tflops'] * 1e12), by * args.batch / (pk['hbm_gbs'] * 1e9)) * 1e3
    print(f'{name:44s} {cls:26s} {n:3d} {ms:8.3f} {100 * ms / tot:5.1f} {fl * args.batch / per / 1e12 if per else 0:8.1f} '
          f'{by * args.batch / per / 1e9 if per else 0:7.0f} {fl / 1e6:10.1f} {by / 1e3:8.1f} {floor:8.3f} '
          f'{ms / floor if floor else 0:7.2f}')
print(f'sum of per-op roofline floors: {floor_tot:.3f} ms ({pk["source"]}); measured {tot:.3f} ms = {tot / floor_tot:.2f}x')
