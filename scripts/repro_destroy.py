import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faulthandler; faulthandler.enable()
import argparse, torch
import bench
mode = sys.argv[1]
args = argparse.Namespace(size='tiny', side=64, joints=8, precision='bf16')
dev = torch.device('cuda', 0)
model = bench.build_model(args, dev)
eng = model.engine(dev)
crops, k = bench.synthetic(4, 64, 0)
if 'host' in mode:
    eng.forward_host(crops.pin_memory(), k.pin_memory())
if 'prof' in mode:
    eng.profile_begin(None); eng.forward(crops.to(dev), k.to(dev)); torch.cuda.synchronize(); eng.profile_end()
if 'plain' in mode:
    eng.forward(crops.to(dev), k.to(dev)); torch.cuda.synchronize()
if 'close' in mode:
    eng.close(); print('closed explicitly', flush=True)
print('exiting', mode, flush=True)
