"""Dev tool: run a long American batch while sampling SM clocks with nvidia-smi."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools.time_variants import american, barrier, timeit
from finite_difference_b200 import DeviceBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
log = open("gpurun_out/clocks_probe.csv", "w")
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                      "--format=csv,noheader", "-lms", "100"], stdout=log)
time.sleep(0.5)
b = american(n, 400, 400, False)
dev = DeviceBatch(b)
for rep in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); dev.solve(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {rep}: {ms:.1f} ms -> {b.node_steps()/ms*1e3:.3e} node-steps/s", flush=True)
time.sleep(0.3)
p.terminate(); log.close()
print(open("gpurun_out/clocks_probe.csv").read())
