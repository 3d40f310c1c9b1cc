export PYTHONPATH=.
o=gpurun_out
bash tools/gpu_time.sh n7 lead0 lead24 lead64 lead160 lead0 lead64 2>&1 | grep -v "^exit"
for v in lead64 lead24; do
  CCG_LIB=$PWD/camera_calibrator_b200/variants/libccg_$v.so timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_intr_sweep -s 2 -c 1 --csv --log-file $o/n7_dram_$v.csv python tools/profile_sweep.py 100000 400 > $o/n7_dram_$v.log 2>&1
  echo "$v ncu exit $?"; grep -E "dram__bytes|gpu__time|hit_rate" $o/n7_dram_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
