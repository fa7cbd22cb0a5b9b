#!/bin/bash
# per-phase timers of swgn_gnss_preprocess (SWGN_GNSS_DEBUG) over 10 epochs of 4096 receivers
mkdir -p gpurun_out
SWGN_GNSS_DEBUG=1 python tools/gnss_epoch_bench.py 4096 10 > gpurun_out/gnss_dbg.json 2> gpurun_out/gnss_dbg.err
python - <<'PY'
import json, re
d = json.load(open('gpurun_out/gnss_dbg.json'))
print(d['ms_per_call'])
lines = [l for l in open('gpurun_out/gnss_dbg.err') if l.startswith('[gnss]')]
calls, cur = [], []
for l in lines:
    m = re.match(r'\[gnss\] (.*?)\s+([0-9.]+) ms', l)
    cur.append((m.group(1).strip(), float(m.group(2))))
    if m.group(1).strip().startswith('pass 2 read-back'):
        calls.append(cur); cur = []
for i, c in enumerate(calls):
    print(i, ' '.join('%s=%.1f' % (k, v) for k, v in c))
PY
