#!/bin/bash
# gpurun --timeout 900 -- tools/gpu_line_gather.sh : line-gather rates + their L2/DRAM traffic (see tools/line_gather.py)
mkdir -p gpurun_out
python tools/line_gather.py 32 > gpurun_out/line_gather.jsonl 2> gpurun_out/line_gather.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum \
    --clock-control none -k regex:km_gather --csv --log-file gpurun_out/line_gather_ncu.csv python tools/line_gather.py 32 > /dev/null 2>&1
cat gpurun_out/line_gather.jsonl
