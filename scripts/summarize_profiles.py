"""Turns the raw evidence brought back in gpurun_out/ (ncu launch list of the bench command, ncu --set full reports, the
head / soft-argmax sweep) into the committed summaries under profiles/.  Usage: python scripts/summarize_profiles.py r1"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'

KEYS = ['ID', 'Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']


def full_summary(rep, out):
    path = os.path.join(G, rep)
    pre = path.replace('.ncu-rep', '.raw.csv')   # raw page exported on the GPU box (the report itself was too big to bring back)
    if os.path.exists(path):
        raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    elif os.path.exists(pre):
        raw = open(pre).read()
    else:
        return
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = [hdr.index(k) for k in KEYS if k in hdr]
    with open(os.path.join(P, out), 'w', newline='') as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in idx])
    print('wrote', out, len(rows) - 2, 'launches')


def launch_list(src, out_csv, out_md):
    path = os.path.join(G, src)
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ik, im, iu, ig = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
    inm = hdr.index('Metric Name')
    per = collections.OrderedDict()
    with open(os.path.join(P, out_csv), 'w', newline='') as f:
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
                mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r[iu], 1)
                a['dram'] += v * mult
    tot = sum(a['us'] for a in per.values())
    with open(os.path.join(P, out_md), 'w') as f:
        f.write(f'| kernel | launches | device time (us) | share | DRAM bytes/launch (MB) |\n|---|---|---|---|---|\n')
        for k, a in sorted(per.items(), key=lambda kv: -kv[1]['us']):
            dram = f"{a['dram'] / a['n'] / 1e6:.1f}" if a.get('dram') else '-'
            f.write(f"| `{k}` | {int(a['n'])} | {a['us']:.1f} | {100 * a['us'] / tot:.1f} % | {dram} |\n")
        f.write(f"\ntotal {tot / 1e3:.2f} ms over {int(sum(a['n'] for a in per.values()))} launches "
                f"(ncu serialises launches and runs them cold-cache: compare SHARES, not absolutes)\n")
    print('wrote', out_csv, out_md)
    # DRAM traffic per launch of each kernel FAMILY (template arguments stripped), read by bench.py for roofline.traffic
    fam = collections.defaultdict(lambda: [0, 0.0])
    for k, a in per.items():
        f_ = k.split('<')[0]
        fam[f_][0] += a['n']
        fam[f_][1] += a.get('dram', 0.0)
    traffic = {k: v[1] / v[0] for k, v in fam.items() if v[1] > 0}
    if traffic:
        with open(os.path.join(P, 'traffic.json'), 'w') as f:
            json.dump(dict(source=f'{out_csv}: mean dram__bytes_read.sum + dram__bytes_write.sum per launch over the bench step '
                                  f'(EfficientNetV2-L@256, 256 crops)', **traffic), f, indent=1)
        print('wrote traffic.json', {k: round(v / 1e6, 1) for k, v in traffic.items()})
    return per


os.makedirs(P, exist_ok=True)
launch_list('bench_launches.csv', f'{tag}_bench_launch_list_ncu.csv', f'{tag}_bench_launch_list_summary.md')
full_summary('tc_conv_r1_final.ncu-rep', f'{tag}_tc_conv_kernel_ncu_full_summary.csv')
full_summary('dwconv_r1.ncu-rep', f'{tag}_dwconv3x3_pool_kernel_ncu_full_summary.csv')
full_summary('softargmax_r1.ncu-rep', f'{tag}_softargmax_and_fused_head_ncu_full_summary.csv')
# captures of the re-entry session: TMA-staged depthwise kernel and the stage-5 projection GEMM (1344 -> 224, one N tile)
full_summary('r1_dw_tma.ncu-rep', f'{tag}_dw3x3s1_tma_kernel_ncu_full_summary.csv')
full_summary('r1_dw_tma_v2.ncu-rep', f'{tag}_dw3x3s1_tma_kernel_ffma2_ncu_full_summary.csv')
full_summary('r1_tc_project.ncu-rep', f'{tag}_tc_conv_kernel_projection_1344x224_ncu_full_summary.csv')
hs = os.path.join(G, 'head_sweep.jsonl')
if os.path.exists(hs):
    with open(hs) as f, open(os.path.join(P, f'{tag}_head_sweep.jsonl'), 'w') as o:
        o.write(f.read())
    print('wrote head sweep')
