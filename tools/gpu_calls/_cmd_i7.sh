export PYTHONPATH=.
timeout 300 python tools/profile_rig.py > gpurun_out/i7_rig_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rig_sweep -s 3 -c 1 -f -o gpurun_out/i7_rig_sweep python tools/profile_rig.py > gpurun_out/i7_rig_ncu.log 2>&1
echo "ncu exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rig_solve -s 3 -c 1 -f -o gpurun_out/i7_rig_solve python tools/profile_rig.py > gpurun_out/i7_rig_ncu2.log 2>&1
echo "ncu exit $?"
