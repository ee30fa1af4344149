"""Where a call of the warp-MMA generation engine spends its time: one record per (warp, job) from the kernel itself
(MMB_MMA_TRACE=1: start / end in globaltimer ns, SM, CTA, warp, jet, key, segment), turned into the number of busy warps over
time, the idle share of the chip and the job durations per key.

    MMB_MMA_HOME=0|1 MMB_MMA_SEGS=1|3 python tools/mma_timeline.py [--jets 4096]
"""
import argparse
import ctypes
import json
import os
import sys

os.environ["MMB_MMA_TRACE"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from multimodal_particles_b200 import _native  # noqa: E402
from multimodal_particles_b200.epic import as_u8  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--jets", type=int, default=4096)
ap.add_argument("--dump", default=None, help="write the raw records (npy)")
args = ap.parse_args()

dev = torch.device("cuda:0")
cfg, model = bench.build_model(dev)
native = model.encoder.native_model(dev)
table = model.step_table()
b = bench.source_batch(args.jets, 1234)
x0, k0, m = b.source_continuous.to(dev).contiguous(), as_u8(b.source_discrete.to(dev)), as_u8(b.source_mask.to(dev))
lib = _native.load()
lib.mmb_debug_read_mma_trace.restype = ctypes.c_longlong
lib.mmb_debug_read_mma_trace.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
words = 4 * (1 << 18)
buf = np.zeros(words, dtype=np.uint64)
ms = []
for i in range(4):
    x, k = x0.clone(), k0.clone()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    native.generate(x, k, m, table, seed=1, precision="f16")
    e.record()
    torch.cuda.synchronize()
    ms.append(s.elapsed_time(e))
n = lib.mmb_debug_read_mma_trace(buf.ctypes.data, words)   # records of the LAST call (the buffer is reset per call)
r = buf[:4 * n].reshape(n, 4)
t0, t1 = r[:, 0].astype(np.int64), r[:, 1].astype(np.int64)
sm = (r[:, 2] >> np.uint64(32)).astype(np.int64)
key = ((r[:, 3] >> np.uint64(8)) & np.uint64(0xff)).astype(np.int64)
seg = (r[:, 3] & np.uint64(0xff)).astype(np.int64)
org = t0.min()
t0, t1 = t0 - org, t1 - org
span = t1.max()
out = {"jets": args.jets, "home": os.environ.get("MMB_MMA_HOME", "default"), "segs": os.environ.get("MMB_MMA_SEGS", "default"),
       "event_ms": [round(v, 4) for v in ms], "records": int(n), "first_start_to_last_end_us": span / 1e3}
warps = 148 * 16
busy = (t1 - t0).sum()
out["busy_share_of_warp_time"] = float(busy / (warps * span))
# busy warps over time, 20 slices
edges = np.linspace(0, span, 21)
prof = []
for a, bb in zip(edges[:-1], edges[1:]):
    ov = np.clip(np.minimum(t1, bb) - np.maximum(t0, a), 0, None).sum()
    prof.append(round(float(ov / ((bb - a) * warps)), 3))
out["busy_warp_fraction_in_20_slices"] = prof
# per SM: when does it end, how busy
ends = np.array([t1[sm == s_].max() if (sm == s_).any() else 0 for s_ in range(148)])
out["sm_end_us_min_med_max"] = [float(np.min(ends) / 1e3), float(np.median(ends) / 1e3), float(np.max(ends) / 1e3)]
dur = {}
for kk in sorted(set(key.tolist())):
    d = (t1 - t0)[key == kk] / 1e3
    dur[int(kk)] = {"jobs": int(d.size), "us_mean": round(float(d.mean()), 1), "us_p10": round(float(np.percentile(d, 10)), 1),
                    "us_p90": round(float(np.percentile(d, 90)), 1)}
out["job_us_by_key"] = dur
# copies of the step loop an SM runs at the same time (keys of its busy warps), sampled
mix = []
for tt in np.linspace(0.05, 0.95, 19) * span:
    live = (t0 <= tt) & (t1 > tt)
    per = [len(set((key[live & (sm == s_)] % 4).tolist())) for s_ in range(0, 148, 4)]
    mix.append(round(float(np.mean(per)), 2))
out["mean_distinct_last_warp_copies_per_sm"] = mix
print(json.dumps(out))
if args.dump:
    np.save(args.dump, r)
