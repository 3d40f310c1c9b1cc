"""Measure the two roofline denominators this path needs on the current GPU: FP64 FMA
peak (not in MEASURED_PEAKS.json) and a device copy bandwidth (cross-check of hbm_gbs)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import camera_calibrator_b200 as cc

ctx = cc.Context(0)
out = dict(device=ctx.device_info(), fp64_tflops=ctx.measure_fp64_peak(),
           copy_gbs=ctx.measure_copy_bandwidth(1 << 30))
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/peaks.json", "w"))
