#!/usr/bin/env python
"""bench.py - crops/s of the MeTRAbs crop-model hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port, torch-cpu)

One JSON line on rank 0.  `value`: whole-job crops/s with the crops already resident in HBM.  `e2e`: the same metric
through the reference-facing host-buffer call (mtb_forward_host: pinned host crops -> H2D -> forward -> D2H joints).
`roofline`: the dominant kernel class, timed live with CUDA events on the launching stream inside the timed region.
`cpu_baseline`: the oracle port on the box's host cores on a bounded sample (rank 0, N=1 only)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'crops/sec'
NAMES = {'s': 'efficientnetv2-s', 'm': 'efficientnetv2-m', 'l': 'efficientnetv2-l', 'tiny': 'efficientnetv2-tiny',
         'resnet50': 'resnet50', 'mobilenetv3-small': 'mobilenetv3-small'}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--size', default='l', choices=list(NAMES))
    ap.add_argument('--side', type=int, default=256)
    ap.add_argument('--stride', type=int, default=32, help='stride_test (ResNet-50 config c2: 8)')
    ap.add_argument('--depth', type=int, default=8, help='heatmap depth D (config c2: 32)')
    ap.add_argument('--joints', type=int, default=24)
    ap.add_argument('--batch', type=int, default=256, help='crops per GPU per step (weak scaling)')
    ap.add_argument('--precision', default=os.environ.get('MTB_BENCH_PRECISION', 'bf16'), choices=['fp32', 'bf16', 'tf32x3'])
    ap.add_argument('--cpu-sample', type=int, default=8, help='crops per CPU-baseline forward')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--graph', type=int, default=int(os.environ.get('MTB_BENCH_GRAPH', '0')),
                    help='1: replay the forward from a CUDA graph in the `value` region (mtb_forward never syncs or allocates)')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p['hbm_gbs'], tflops=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    tflops_burst=p['bf16_tflops'], source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, tflops=1400.0, tflops_burst=1590.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '100'], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        window = [ln for (t, ln) in self.lines if t0 is None or (t0 <= t <= t1 + 0.15)]
        if not window:  # timed region shorter than the sampling period: use the samples taken under warm-up load
            window = [ln for (_, ln) in self.lines[-3:]]
        for ln in window:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[5:9]):
                if v.lower() == 'active':
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def build_model(args, device):
    import metrabs_b200
    from metrabs_b200.backbones import efficientnet as E
    from metrabs_b200.init import conditioned_random_init_
    from metrabs_b200.models.metrabs import Metrabs
    import types
    metrabs_b200.set_config(metrabs_b200.Config(proc_side=args.side, precision=args.precision,
                                                stride_test=getattr(args, 'stride', 32), depth=getattr(args, 'depth', 8)))
    ji = types.SimpleNamespace(names=[f'j{i}' for i in range(args.joints)], stick_figure_edges=[(0, 1)],
                               n_joints=args.joints)
    if args.size == 'resnet50':
        from metrabs_b200.backbones import resnet
        backbone = resnet.resnet50()
    elif args.size == 'mobilenetv3-small':
        from metrabs_b200.backbones import mobilenet_v3
        backbone = mobilenet_v3.mobilenet_v3_small()
    else:
        backbone = torch.nn.Sequential(E.PreprocLayer(), E.EfficientNet(args.size).features)
    model = Metrabs(backbone, ji).eval()
    conditioned_random_init_(model, seed=0)
    return model.to(device) if device is not None else model


def synthetic(batch, side, seed):
    g = torch.Generator().manual_seed(seed)
    crops = torch.rand(batch, 3, side, side, generator=g)
    f = 1000 + 500 * torch.rand(batch, generator=g)
    k = torch.zeros(batch, 3, 3)
    k[:, 0, 0] = f
    k[:, 1, 1] = f
    k[:, 0, 2] = side / 2
    k[:, 1, 2] = side / 2
    k[:, 2, 2] = 1
    return crops, k


def effective_cores():
    """Host cores this process may actually use: affinity mask capped by the cgroup CPU quota (a container that
    reports 128 logical CPUs but is throttled to a few makes torch-cpu oversubscribe badly)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:
            quota, period = f.read().split()
        if quota != 'max':
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        try:
            with open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us') as f:
                quota = int(f.read())
            with open('/sys/fs/cgroup/cpu/cpu.cfs_period_us') as f:
                period = int(f.read())
            if quota > 0:
                n = max(1, min(n, int(quota / period + 0.5)))
        except Exception:
            pass
    return n


def best_thread_count():
    """All the host threads torch-cpu can actually USE: probes a small forward at a few thread counts and keeps the
    fastest (on shared hosts 'all logical CPUs' can be 100x slower than a moderate count)."""
    import argparse as _ap
    from oracle import port
    eff = effective_cores()
    cands = sorted({eff, min(eff, 64), min(eff, 32), min(eff, 16), min(eff, 8)}, reverse=True)
    pa = _ap.Namespace(size='s', side=128, joints=24, precision='fp32')
    model = build_model(pa, None)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    pcfg = port.PathConfig(proc_side=128)
    spec = port.effnet_spec(NAMES['s'])
    crops, k = synthetic(2, 128, 0)
    best, best_t = cands[-1], float('inf')
    with torch.inference_mode():
        for n in cands:
            torch.set_num_threads(n)
            port.metrabs_forward(sd, spec, pcfg, 24, crops, k)
            t0 = time.perf_counter()
            port.metrabs_forward(sd, spec, pcfg, 24, crops, k)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = n, dt
    return best


def cpu_reference_forward(args, n_crops, iters, warmup):
    """The reference's CPU path (oracle port, torch-cpu fp32, all host threads) on `n_crops` synthetic crops."""
    from oracle import port
    torch.set_num_threads(best_thread_count())
    model = build_model(args, None)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    pcfg = port.PathConfig(proc_side=args.side, stride_test=getattr(args, 'stride', 32), depth=getattr(args, 'depth', 8))
    if args.size in ('resnet50', 'mobilenetv3-small'):
        from oracle import port_tf_backbones as tfb
        spec = tfb.ResNet50Spec(pcfg) if args.size == 'resnet50' else tfb.MobileNetV3SmallSpec(pcfg)
    else:
        spec = port.effnet_spec(NAMES[args.size])
    crops, k = synthetic(n_crops, args.side, 0)
    times = []
    with torch.inference_mode():
        for i in range(warmup + iters):
            t0 = time.perf_counter()
            port.metrabs_forward(sd, spec, pcfg, args.joints, crops, k)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return n_crops / (sum(times) / len(times)), torch.get_num_threads(), sum(times) / len(times)


def workload_name(args):
    return (f'{NAMES[args.size]} {args.side}x{args.side} J={args.joints} D={args.depth} stride={args.stride}, {args.batch} crops/GPU/step '
            f'(BASELINE.json metric: crops/sec, 256x256, EffNetV2-L, 24 joints)')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.cpu_sample
    v, cores, sec = cpu_reference_forward(args, n, max(args.steps, 1), min(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'crops/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'note': 'reference arm = oracle port of the reference forward on '
                   'torch-cpu (the reference is pure Python; /root/reference is absent on the GPU box)'},
        'cpu_baseline': {'value': v, 'unit': 'crops/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{n} crops per step, {args.steps} steps'},
        'e2e': {'value': v, 'unit': 'crops/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torch.distributed.run --nproc-per-node N for --gpus N')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    model = build_model(args, device)
    eng = model.engine(device)
    B, S, J = args.batch, args.side, args.joints
    if world > 1:
        def bcast(raw):
            t = torch.tensor(list(raw) if raw is not None else [0] * 128, dtype=torch.uint8, device=device)
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        eng.comm_init(rank, world, bcast)
    crops_h, k_h = synthetic(B, S, 100 + rank)
    crops_h, k_h = crops_h.pin_memory(), k_h.pin_memory()
    out_h = torch.empty(B, J, 3).pin_memory()
    crops_d, k_d = crops_h.to(device), k_h.to(device)
    out_d = torch.empty(B, J, 3, device=device)

    def step():
        eng.forward(crops_d, k_d, out=out_d)
        if world > 1:
            return eng.allgather(out_d)
        return out_d

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graph = None
    if args.graph and world == 1:
        step()
        torch.cuda.synchronize()
        graph = eng.capture_forward(crops_d, k_d, out_d)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # warm-up; the first one with every kernel class bracketed by events to find the dominant class
    eng.profile_begin(None)
    step()
    torch.cuda.synchronize()
    prof_all = eng.profile_end()
    for _ in range(max(args.warmup - 1, 0)):
        step()
    dom_name = max(prof_all, key=lambda n: prof_all[n]['ms'])
    dom_cls = prof_all[dom_name]['cls']
    total_ms_all = sum(v['ms'] for v in prof_all.values())

    barrier()
    graph_ms = None
    if graph is not None:  # CUDA-graph replay of the same K steps (no per-kernel events inside): reported beside `value`
        for _ in range(2):
            graph.replay()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(args.steps):
            graph.replay()
        g1.record()
        torch.cuda.synchronize()
        graph_ms = g0.elapsed_time(g1)
    eng.profile_begin([dom_cls])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = sampler.mark()
    ev0.record()
    launches = 0
    for _ in range(args.steps):
        step()
        launches += eng.last_launch_count + (1 if world > 1 else 0)
    ev1.record()
    barrier()
    t_end = sampler.mark()
    elapsed_ms = ev0.elapsed_time(ev1)
    prof_dom = eng.profile_end()[dom_name]
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None

    # end-to-end through the host-buffer entry points (pinned host crops in, host joints out, EVERY step).  Back-to-back
    # batches go through the pipelined pair mtb_forward_host_submit / _wait (two slots): the H2D copy of step i+1 runs on
    # the library's copy stream while step i computes; every step still copies its own inputs and reads its own joints.
    # Its results must equal the synchronous call's bit for bit, otherwise (or on any error) the synchronous loop is timed.
    eng.forward_host(crops_h, k_h, out_h)
    ref_out = out_h.clone()
    e2e_mode = 'synchronous mtb_forward_host per step'
    e2e_ms = None
    pipe_ok = not os.environ.get('MTB_BENCH_SYNC_E2E')
    pipe_ms = None
    if pipe_ok:
        out_hs = [torch.empty(B, J, 3).pin_memory(), torch.empty(B, J, 3).pin_memory()]
        try:
            for s_ in (0, 1):  # warm both slots (staging allocations)
                eng.forward_host_submit(crops_h, k_h, out_hs[s_], s_)
                eng.forward_host_wait(s_)
        except Exception as e:  # noqa: BLE001
            print(f'bench: pipelined host path failed ({e!r})', file=sys.stderr)
            pipe_ok = False
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            s_ = i & 1
            if pipe_ok:
                try:
                    eng.forward_host_wait(s_)  # the step that used this slot two steps ago has delivered its joints
                    eng.forward_host_submit(crops_h, k_h, out_hs[s_], s_)
                except Exception as e:  # noqa: BLE001
                    print(f'bench: pipelined host path failed ({e!r})', file=sys.stderr)
                    pipe_ok = False
            if world > 1:  # every rank issues the same number of collectives whatever happened above
                eng.allgather(out_d)
        if pipe_ok:
            try:
                eng.forward_host_wait(0)
                eng.forward_host_wait(1)
            except Exception as e:  # noqa: BLE001
                print(f'bench: pipelined host path failed ({e!r})', file=sys.stderr)
                pipe_ok = False
        barrier()
        pipe_ms = (time.perf_counter() - t0) * 1e3
        if pipe_ok and not (torch.equal(out_hs[0], ref_out) and torch.equal(out_hs[1], ref_out)):
            print('bench: pipelined host path disagrees with mtb_forward_host', file=sys.stderr)
            pipe_ok = False
    if world > 1:  # the fallback decision is collective: either every rank reports the pipelined loop or none does
        flag = torch.tensor([1 if pipe_ok else 0], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        pipe_ok = bool(flag.item())
    if pipe_ok:
        e2e_ms = pipe_ms
        e2e_mode = 'pipelined mtb_forward_host_submit/_wait, 2 slots (H2D of step i+1 overlaps the forward of step i)'
    else:
        print('bench: timing the synchronous mtb_forward_host loop', file=sys.stderr)
    if e2e_ms is None:
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.forward_host(crops_h, k_h, out_h)
            if world > 1:
                eng.allgather(out_d)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = world * B * args.steps / (elapsed_ms / 1e3)
    e2e = world * B * args.steps / (e2e_ms / 1e3)
    tensor_bound = prof_dom['flops'] > 0 and (prof_dom['flops'] / max(prof_dom['bytes'], 1)) > 100
    if tensor_bound:
        achieved = prof_dom['flops'] / (prof_dom['ms'] / 1e3) / 1e12
        peak, unit, bound = pk['tflops'], 'TFLOP/s', 'tensor'
    else:
        achieved = prof_dom['bytes'] / (prof_dom['ms'] / 1e3) / 1e9
        peak, unit, bound = pk['hbm_gbs'], 'GB/s', 'hbm'
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom_name)
    flops_crop = eng.backbone_flops_per_crop
    line = {
        'metric': METRIC, 'value': value, 'unit': 'crops/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': {'bf16': 'bf16', 'fp32': 'f32', 'tf32x3': 'tf32x3'}[args.precision], 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'global_batch': world * B, 'parallelism': f'dp{world}',
                   'precision_mode': args.precision, 'weights': 'conditioned random init (metrabs_b200/init.py)',
                   'l2_policy': f'inputs larger than L2: {B * 3 * S * S * 4 / 1e6:.0f} MB of crops per step',
                   'backbone_gflop_per_crop': flops_crop / 1e9,
                   'tensor_util_of_peak': value / world * flops_crop / 1e12 / pk['tflops'],
                   'peaks': pk['source'],
                   'cuda_graph_replay_crops_per_s': (world * B * args.steps / (graph_ms / 1e3)) if graph_ms else None},
        'e2e': {'value': e2e, 'unit': 'crops/s', 'h2d_bytes_per_step': B * 3 * S * S * 4 + B * 36,
                'd2h_bytes_per_step': B * J * 3 * 4, 'mode': e2e_mode},
        'gpu_launches': launches,
        'clocks': clocks,
        'roofline': {'kernel': dom_name, 'bound': bound, 'achieved': achieved, 'peak': peak, 'unit': unit,
                     'frac': achieved / peak, 'traffic': traffic,
                     'launches_timed': prof_dom['launches'], 'avg_launch_us': prof_dom['ms'] * 1e3 / prof_dom['launches'],
                     'share_of_step': prof_all[dom_name]['ms'] / total_ms_all,
                     'class_ms_first_step': {n: round(v['ms'], 3) for n, v in prof_all.items()}},
    }
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sec = cpu_reference_forward(args, args.cpu_sample, 2, 1)
        line['cpu_baseline'] = {'value': v, 'unit': 'crops/s', 'cores': cores, 'kind': 'port',
                                'sample': f'{args.cpu_sample} crops x 2 timed forwards of the oracle port '
                                          f'(torch-cpu fp32), {sec:.1f} s each'}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    import faulthandler
    faulthandler.enable()
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
