"""Per-phase device timeline of one config (debug/measurement aid): python tools/trace_window.py c2 [batch]
Writes gpurun_out/trace_<cfg>.json.  Phase ids: 10 init, 11 W1 (CSR build), 13 W2+W3, 14 W4+round 1, 1 PROP, 2 GREEDY, 3 FORCE,
4 D1, 5 D2, 6/7 EVAL."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ms_slam_b200 import msgen
from ms_slam_b200.engine import Engine, DeviceView
from ms_slam_b200.window import pack_view

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
soa = "--soa" in sys.argv          # default: the transport form bench.py ships (discovery order, packed layout)
views = [msgen.make_config(name, s)[0] for s in range(batch)]
if not soa:
    views = [pack_view(v.compact().discovery_order()) for v in views]
N = msgen.CONFIGS[name]["N"]
eng = Engine(N=N, lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA)
dv = [DeviceView(eng, v) for v in views]
for _ in range(3):
    eng.solve_batch(dv)
eng.trace(True)
eng.solve_batch(dv)
st = eng.stats()
ends = []
for i in range(batch):
    t = eng.get_trace(i)
    ends.append(t[-1][2] / 1e3 if t else -1)
print("per-window end (us):", " ".join(f"{e:.0f}" for e in ends))
tr = eng.get_trace(0)
names = {10: "init", 11: "W1", 12: "W2", 13: "W3", 15: "W4pairs", 14: "W4R1", 1: "PROP", 2: "GREEDY", 3: "FORCE", 4: "D1", 5: "D2", 6: "EVAL", 7: "EVALV", 8: "TAIL", 9: "BOUND", 20: "t-gather", 21: "t-PROP", 22: "t-GREEDY", 30: "t-rows", 31: "t-maxn"}
prev = 0
out = []
for ph, free, ns in tr:
    out.append(dict(phase=names.get(ph, ph), free=free, t_us=ns / 1e3, dt_us=(ns - prev) / 1e3))
    prev = ns
print(f"{name} x{batch}: kernel {st['last_device_ms']*1e3:.1f} us, grid {st['grid_ctas']}")
for o in out:
    print(f"  {str(o['phase']):7s} free={o['free']:7d}  t={o['t_us']:9.1f}  dt={o['dt_us']:8.1f}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(config=name, batch=batch, kernel_us=st['last_device_ms'] * 1e3, grid=st['grid_ctas'], phases=out),
          open(os.path.join(ROOT, "gpurun_out", f"trace_{name}_{batch}.json"), "w"), indent=1)
