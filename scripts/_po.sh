python scripts/phase_overhead.py 192 0; python scripts/phase_overhead.py 192 1
for ov in 0 1; do ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/po_$ov.csv python scripts/phase_overhead.py 192 $ov > /dev/null 2>&1; python scripts/kernel_times.py gpurun_out/po_$ov.csv | grep -v "at::" ; done
