#!/bin/bash
# dev helper (GPU box): forward path comparison tma vs lazy on cfg2 and cfg5
for p in tma lazy; do
  echo "== GG_FWD_PATH=$p cfg2"; GG_FWD_PATH=$p python bench.py --steps 40 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k={x['kernel']:x['ms'] for x in d['roofline']['kernels']}
print('views/s %.1f' % d['value'], {a:round(b,4) for a,b in k.items()})"
done
for p in tma lazy; do
  echo "== GG_FWD_PATH=$p cfg5"; GG_FWD_PATH=$p timeout 300 python tools/run_configs.py --config cfg5 --views 2 --steps 1 2>&1 | tail -1 | cut -c1-700
done
