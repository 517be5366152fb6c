#!/bin/bash
# dev helper: compare libgg_raster.so build variants (GG_RASTER_LIB override) on the GPU box
for lib in "$@"; do
  GG_RASTER_LIB=$PWD/$lib python bench.py --steps 40 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k={x['kernel']:x['ms'] for x in d['roofline']['kernels']}
print('$lib', 'views/s %.1f' % d['value'], 'bwd %.4f fwd %.4f sort %.4f' % (k['blend_bwd'], k['blend_fwd'], k['sort_pack']))"
done
