#!/bin/bash
# GPU run r02f (2 GPUs): ncu of the multi-GPU level-1 kernel (k_collect_owned) in the single-process form
# (dnagpu_bench --host --gpus 2 = dnagpu_create_multi), with the NVLink / peer-aperture counters next to --set full.
cd "$(dirname "$0")/.."
O=gpurun_out
X="./dna-sequences-pg-extension_b200/dnagpu_bench --bases 1000000000 --k 31 --seed 5 --steps 1 --host --gpus 2"
ncu --query-metrics 2>/dev/null | grep -iE "nvl|peer|aperture" | head -60 > $O/r02f_nvlink_metric_names.txt
$X > $O/r02f_cbench.json 2> $O/r02f_cbench.err; echo "plain rc=$?" > $O/r02f_status.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_collect_owned" -s 2 -c 2 -f -o $O/r02f_collect_owned_full $X > $O/r02f_ncu1.log 2>&1; echo "ncu full rc=$?" >> $O/r02f_status.txt
M=$(grep -oE "^(nvlrx__bytes|nvltx__bytes|lts__t_sectors_aperture_peer|lts__t_sectors_srcunit_tex_aperture_peer|lts__t_bytes_aperture_peer|l1tex__m_xbar2l1tex_read_sectors_mem_lg_aperture_peer)[a-z_.]*" $O/r02f_nvlink_metric_names.txt | sort -u | tr '\n' ',' | sed 's/,$//')
[ -n "$M" ] && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,$M --clock-control none -k regex:"k_collect_owned" -s 2 -c 2 --csv \
   --log-file $O/r02f_collect_owned_nvlink.csv $X > $O/r02f_ncu2.log 2>&1; echo "ncu nvlink rc=$? metrics=$M" >> $O/r02f_status.txt
cat $O/r02f_status.txt; head -30 $O/r02f_nvlink_metric_names.txt
