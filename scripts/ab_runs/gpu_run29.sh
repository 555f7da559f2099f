echo "=== trace 32x32 full"; MTB_TC_DEBUG=32 MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 1 2>&1 | head -4 > gpurun_out/trace_full.txt; python - <<'PY'
import re
for line in open('gpurun_out/trace_full.txt'):
    if 'mma(' in line:
        v=[int(x.strip('[]')) for x in line.split(':')[1].split()]
        starts=v[0::2]; ends=v[1::2]
        print('tiles traced', len(starts))
        print('tile starts', starts[:6], '...', starts[-6:])
        d=[b-a for a,b in zip(starts[:-1],starts[1:])]
        print('periods first 10', d[:10]); print('periods last 20', d[-20:])
        print('issue times last 10', [e-s for s,e in zip(starts[-10:],ends[-10:])])
    if 'epilogue' in line:
        v=[int(x) for x in line.split(':')[1].split()]
        print('epi last', v[-12:])
    if 'producer' in line:
        v=[int(x) for x in line.split(':')[1].split()]
        print('loader last', v[-12:])
PY
