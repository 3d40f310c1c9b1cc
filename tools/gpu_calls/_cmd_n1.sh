export PYTHONPATH=.
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -x -q -m gpu > $o/n1_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $o/n1_pytest.log
bash tools/gpu_time.sh n1 head noldsm ldsm ldsm_j96 head ldsm
