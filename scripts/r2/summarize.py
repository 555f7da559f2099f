"""Turns the round-2 final evidence (gpurun_out/r2_final_*, scripts/r2/gpu_final2.sh) into the committed summaries under
profiles/: bench lines, launch lists with per-kernel DRAM bytes, traffic.json keyed by (workload, precision) for bench.py's
roofline.traffic, ncu --set full summaries, per-op profiles.  Usage: python scripts/r2/summarize.py"""
import collections
import csv
import glob
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
CLASS_OF = {'tc_conv_kernel': 'tc_conv_kernel', 'fmb_kernel': 'fmb_kernel', 'tc32_conv_kernel': 'tc32_conv_kernel',
            'se_scale_kernel': 'se_scale_kernel', 'stem3x3s2_kernel': 'stem_conv_kernel', 'stem_conv_wide_kernel': 'stem_conv_kernel',
            'dw3x3s1_tma_kernel': 'dwconv_kernel'}


def launch_list(src, tag):
    path = os.path.join(G, src)
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ik, im, iu, ig, inm = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Grid Size', 'Metric Name'))
    per = collections.OrderedDict()
    with open(os.path.join(P, f'r2_bench_launch_list_{tag}_ncu.csv'), 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['launch', 'kernel', 'grid', 'metric', 'unit', 'value'])
        for r in data:
            if len(r) <= im:
                continue
            w.writerow([r[0], r[ik], r[ig], r[inm], r[iu], r[im]])
            name = r[ik].split('(')[0].replace('void ', '')
            v = float(r[im].replace(',', ''))
            a = per.setdefault(name, collections.defaultdict(float))
            if r[inm] == 'gpu__time_duration.sum':
                a['n'] += 1
                a['us'] += v / 1e3 if r[iu] == 'ns' else (v * 1e3 if r[iu] == 'ms' else v)
            elif r[inm].startswith('dram__bytes'):
                a['dram'] += v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r[iu], 1)
    tot = sum(a['us'] for a in per.values())
    with open(os.path.join(P, f'r2_bench_launch_list_{tag}_summary.md'), 'w') as f:
        f.write('| kernel | launches | device time (us) | share | DRAM bytes/launch (MB) |\n|---|---|---|---|---|\n')
        for k, a in sorted(per.items(), key=lambda kv: -kv[1]['us']):
            dram = f"{a['dram'] / a['n'] / 1e6:.1f}" if a.get('dram') else '-'
            f.write(f"| `{k}` | {int(a['n'])} | {a['us']:.1f} | {100 * a['us'] / tot:.1f} % | {dram} |\n")
        f.write(f"\ntotal {tot / 1e3:.2f} ms over {int(sum(a['n'] for a in per.values()))} launches "
                f"(ncu serialises launches and runs them cold-cache: compare SHARES, not absolutes)\n")
    fam = collections.defaultdict(lambda: [0, 0.0])
    for k, a in per.items():
        f_ = k.split('<')[0]
        fam[f_][0] += a['n']
        fam[f_][1] += a.get('dram', 0.0)
    out = {}
    for k, v in fam.items():
        if v[1] > 0:
            out[CLASS_OF.get(k, k)] = out.get(CLASS_OF.get(k, k), 0.0)
    # per-launch mean over the kernels that map to one bench class
    acc = collections.defaultdict(lambda: [0, 0.0])
    for k, v in fam.items():
        c = CLASS_OF.get(k, k)
        acc[c][0] += v[0]
        acc[c][1] += v[1]
    return {c: v[1] / v[0] for c, v in acc.items() if v[1] > 0}


os.makedirs(os.path.join(P, 'r2_bench_lines'), exist_ok=True)
for f in glob.glob(os.path.join(G, 'r2_final_bench_*.json')):
    shutil.copy(f, os.path.join(P, 'r2_bench_lines', os.path.basename(f).replace('r2_final_', '')))
traffic = {'source': 'profiles/r2_bench_launch_list_<precision>_ncu.csv: mean dram__bytes_read.sum + dram__bytes_write.sum per launch of '
                     'every kernel class over one warm bench step, keyed by (workload, precision)'}
for tag in ('bf16', 'tf32x3'):
    t = launch_list(f'r2_final_launches_{tag}.csv', tag)
    if t:
        traffic[f'l@256:b256:{tag}'] = t
        print(tag, {k: round(v / 1e6, 1) for k, v in t.items()})
if len(traffic) > 1:
    with open(os.path.join(P, 'traffic.json'), 'w') as f:
        json.dump(traffic, f, indent=1)
for r, out in (('tc_conv', 'r2_tc_conv_kernel_final_ncu_full_summary.csv'), ('fmb', 'r2_fmb_kernel_pair_ncu_full_summary.csv'),
               ('tc32_conv', 'r2_tc32_conv_kernel_final_ncu_full_summary.csv'), ('dw', 'r2_dw3x3s1_tma_kernel_ncu_full_summary.csv')):
    src = os.path.join(G, f'r2_final_{r}.summary.csv')
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, out))
for src, out in (('r2_final_op_profile_bf16_b256.txt', 'r2_op_profile_bf16_b256.txt'), ('r2_final_op_profile_tf32x3_b256.txt', 'r2_op_profile_tf32x3_b256.txt'),
                 ('r2_final_head_sweep.jsonl', 'r2_head_sweep.jsonl'), ('r2_final_smoke.log', 'r2_smoke.log')):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, out))
log = os.path.join(G, 'r2_final_gpu_tests.log')
if os.path.exists(log):
    with open(log) as f, open(os.path.join(P, 'r2_gpu_tests_tail.txt'), 'w') as o:
        o.write(''.join(f.readlines()[-8:]))
print('done')
