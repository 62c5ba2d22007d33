#!/bin/bash
# tools/profile_round.sh -- the profiling pass of a round on ONE GPU (run under gpurun): tests, bench, ncu launch list, DRAM traffic
# of the packed launches, full-set captures of the streaming scorer (global, glocal) and of the packed traceback kernel.
# Everything lands in gpurun_out/; the summaries worth keeping are copied to profiles/ by hand (names per round).
R=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_gputests.log; tail -2 gpurun_out/${R}_gputests.log
python bench.py > gpurun_out/${R}_bench_n1_final.json 2> gpurun_out/${R}_bench_n1_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err; echo "reference arm rc=$?"
# (each ncu pass only after the same command has exited 0 without ncu: the bench above)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --no-cpu-baseline --merge-pairs 100000 --no-pipeline --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mm16g_kernel -c 12 --csv --log-file gpurun_out/${R}_ncu_dram_traffic_mm16g.csv \
    python bench.py --no-cpu-baseline --merge-pairs 0 --no-pipeline --no-glocal --steps 1 --warmup 1 > gpurun_out/ncu_traffic.log 2>&1; echo "traffic rc=$?"
# the reports with source are tens of MB each and gpurun_out/ travels back only below 64 MiB: export what is read, drop the report
export_report() {   # $1 = report stem
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > /tmp/$1_source.csv 2>/dev/null && python tools/ncu_source_table.py /tmp/$1_source.csv > gpurun_out/$1_source_table.txt 2>&1
  rm -f gpurun_out/$1.ncu-rep
}
python tools/quick_mm.py --queries 6000 > gpurun_out/quick_global.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mm16g_kernel -s 5 -c 1 -o gpurun_out/${R}_mm16g_global_final -f python tools/quick_mm.py --queries 6000 --steps 1 > gpurun_out/ncu_g.log 2>&1; echo "global capture rc=$?"
export_report ${R}_mm16g_global_final
python tools/quick_mm.py --queries 6000 --mode glocal > gpurun_out/quick_glocal.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mm16g_kernel -s 5 -c 1 -o gpurun_out/${R}_mm16g_glocal_final -f python tools/quick_mm.py --queries 6000 --steps 1 --mode glocal > gpurun_out/ncu_gl.log 2>&1; echo "glocal capture rc=$?"
export_report ${R}_mm16g_glocal_final
python tools/probe_pairs.py 300000 > gpurun_out/probe_pairs.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tb16_kernel -s 1 -c 1 -o gpurun_out/${R}_tb16_final -f python tools/probe_pairs.py 300000 > gpurun_out/ncu_tb.log 2>&1; echo "tb16 capture rc=$?"
export_report ${R}_tb16_final
du -sh gpurun_out
