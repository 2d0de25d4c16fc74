#!/bin/bash
# quick env-variable sweep of the single-GPU step (config 2, Default): sort interval, tile shape, trail chunk height
run() { echo -n "$* : "; env "$@" timeout 300 python bench.py --steps 600 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernels']; print('%.4e  %.1f us/step  agents %.1f trail %.1f sort %.1f' % (d['value'], d['ms_per_step']*1e3, k['agents']['ms']*1e3, k['trail']['ms']*1e3, k['sort_ms_per_step']*1e3))"; }
run SM_SORT_INTERVAL=16
run SM_SORT_INTERVAL=24
run SM_SORT_INTERVAL=32
run SM_SORT_INTERVAL=48
run SM_TILE_SHIFT_X=4 SM_TILE_SHIFT_Y=3
run SM_TILE_SHIFT_X=3 SM_TILE_SHIFT_Y=2
run SM_TILE_SHIFT_X=4 SM_TILE_SHIFT_Y=2
run SM_TILE_SHIFT_X=2 SM_TILE_SHIFT_Y=2
run SM_TRAIL_ROWS_PER_CHUNK=4
run SM_TRAIL_ROWS_PER_CHUNK=16
run SM_TRAIL_ROWS_PER_CHUNK=32
