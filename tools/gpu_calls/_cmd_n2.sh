export PYTHONPATH=.
bash tools/gpu_time.sh n2 ldsm_j96 ldsm_j88 ldsm_j80 ldsm_j96_tail8 ldsm_j96_ad8 ldsm_j96
