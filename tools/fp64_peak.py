"""Record the FP64 roofline denominator with its clocks: profiles/fp64_peak.json (MEASURED_PEAKS.json has no FP64 entry).
    python tools/fp64_peak.py > gpurun_out/fp64_peak.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import ClockSampler
from finite_difference_b200 import _lib

torch.cuda.init()
name = torch.cuda.get_device_name(0)
_lib.fp64_peak_tflops()
runs = []
with ClockSampler(0) as clocks:
    t_end = time.time() + 2.0
    while time.time() < t_end:
        tf, ms = _lib.fp64_peak_tflops()
        runs.append(tf)
sms = torch.cuda.get_device_properties(0).multi_processor_count
c = clocks.summary()
print(json.dumps({"gpu_name": name, "sms": sms, "fp64_tflops_best": max(runs), "fp64_tflops_median": sorted(runs)[len(runs) // 2],
                  "runs": len(runs), "how": "fdm_fp64_peak: 8 independent register-resident DFMA chains per thread, 256 threads x 8 CTAs per SM, "
                  "4096 x 64 FMAs per thread, best of 5 launches per call, CUDA events; repeated for 2 s",
                  "nominal_tflops": sms * 64 * 2 * (c["sm_max_mhz"] or 0) / 1e6, "clocks": c,
                  "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}))
