#!/bin/bash
mkdir -p gpurun_out
for g in 0 32; do
  python tools/gather_modes.py $g > gpurun_out/gather_modes_g$g.json 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,dram__sectors_read.sum \
     --clock-control none -k regex:km_gather --csv --log-file gpurun_out/gather_modes_ncu_g$g.csv python tools/gather_modes.py $g > /dev/null 2>&1
done
python tools/gather_roofline.py gpurun_out/gather_roofline.json > gpurun_out/gather_roofline.log 2>&1
