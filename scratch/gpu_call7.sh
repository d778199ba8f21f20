#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.load(open('$1'));k=d['kernels']
print('$2', round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,2),'Mpts/s | knn query', round(k['pob_knn_grid_query']['ms_per_step'],3), 'build', round(k['pob_knn_grid_build']['ms_per_step'],3), 'fps', round(k['pob_farthest_point_sampling']['ms_per_step'],3), '| cfg1 knn us', round(list(d['ops_cfg1']['ops'].values())[0]['us'],1))"; }
B="python bench.py --no-cpu-baseline --steps 200"
for c in 2.0 3.0 4.0 6.0 1.5; do
  POINTOPS_B200_CELL_PTS=$c timeout 200 $B > gpurun_out/c7_cell$c.json 2>/dev/null; show gpurun_out/c7_cell$c.json "cell_pts $c"
done
