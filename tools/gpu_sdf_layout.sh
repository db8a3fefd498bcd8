#!/bin/bash
# SDF layout experiment ("SDF staging"): linear x-fastest grid against 4 x 4 x 2 bricks, on C5 (512^3 grid, beyond
# L2) and C3: in-loop and isolated / L2-flushed span of the state kernel, then ncu counters of the kernel with cold caches
# (ncu's default flush) and with the loop's warm L2 (--cache-control none).   usage: tools/gpu_sdf_layout.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
out=gpurun_out/sdf_layout_$tag.txt; : > $out
M="gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"
for wl in c5 c3; do
  for layout in linear brick; do
    echo "=== $wl layout=$layout" | tee -a $out
    STOMP_B200_SDF_LAYOUT=$layout timeout 300 python tools/timeline.py $wl 30 2>&1 | grep -E "per iteration|cost  " | tee -a $out
    STOMP_B200_SDF_LAYOUT=$layout timeout 300 python tools/timeline.py $wl 20 flush 2>&1 | grep -E "per iteration|cost  " | tee -a $out
    for cc in all none; do
      STOMP_B200_SDF_LAYOUT=$layout STOMP_B200_GRAPH=0 timeout 600 ncu --metrics $M --clock-control none --cache-control $cc -k regex:states_specialised -s 6 -c 2 --csv --log-file gpurun_out/ncu_sdf_${wl}_${layout}_${cc}_$tag.csv python bench.py --workload $wl --steps 4 --warmup 3 --skip-cpu-baseline --skip-c4 --l2 keep > /dev/null 2>&1
      echo "--- ncu cache-control=$cc" | tee -a $out
      python - gpurun_out/ncu_sdf_${wl}_${layout}_${cc}_$tag.csv <<'PY' | tee -a $out
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if 'Metric Name' in r)
ni, vi, ii, gi = hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID'), hdr.index('Grid Size')
by = collections.defaultdict(dict)
for r in rows:
    if r is hdr or r[ni] == 'Metric Name': continue
    try: by[(r[ii], r[gi])][r[ni]] = float(r[vi].replace(',', ''))
    except ValueError: pass
for (i, g), m in by.items():
    req = m.get('l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 0)
    print(f"  launch {i} grid {g}: {m.get('gpu__time_duration.sum', 0) / 1e3:7.1f} us  dram_read {m.get('dram__bytes_read.sum', 0) / 1e6:8.2f} MB  L2 hit {m.get('lts__t_sector_hit_rate.pct', 0):5.1f}%  L1 hit {m.get('l1tex__t_sector_hit_rate.pct', 0):5.1f}%  "
          f"sectors/request {m.get('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 0) / max(req, 1):5.2f}  L2 read sectors {m.get('lts__t_sectors_srcunit_tex_op_read.sum', 0) / 1e6:7.2f} M  warp instr {m.get('smsp__inst_executed.sum', 0) / 1e6:6.2f} M  long_sb/issue {m.get('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 0):5.2f}")
PY
    done
  done
done
