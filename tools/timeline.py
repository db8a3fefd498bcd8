"""In-pipeline kernel timeline of the STOMP loop (first-CTA-start / last-CTA-end %globaltimer stamps):
    python tools/timeline.py [workload] [iterations]
Prints the median duration of every kernel and the gaps between them, inside the running loop."""
import sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bench
from motion_planners_b200 import binding

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
pb = bench.make_problem(name)
import os
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
shard_mode = 1 if bench.WORKLOADS[name]["kind"] == "batch" else 0
e = binding.engine_for_problem(pb, device=local, world_size=world, rank=rank, shard_mode=shard_mode)
if world > 1:      # under torchrun: rollout (or query) sharding over the ranks; rank 0 prints its own timeline
    _, _, _, dist = bench.dist_setup(world)
    if shard_mode == 0:
        uid = binding.comm_unique_id() if rank == 0 else bytes(128)
        e.comm_init(bench.broadcast_bytes(dist, uid, local))
    print_ = print
    def print(*a, **k):
        if rank == 0:
            print_(*a, **k)
    print(f"{world} ranks, exchange: {e.exchange_kind()}")
e.begin_solve()
e.run(0, 5)
e.set_timeline(True)
flush = len(sys.argv) > 3 and sys.argv[3] == "flush"      # isolated iterations on a cold L2, as bench.py times `value`
if flush:
    fl = bench.L2Flusher(local)
    ms = 0.0
    for i in range(iters):
        fl.flush()
        e.timer_begin()
        e.run(5 + i, 1)
        ms += e.timer_end()
else:
    e.timer_begin()
    e.run(5, iters)
    ms = e.timer_end()
tl = e.timeline(iters)
names = ["sample", "cost", "weights", "update", "apply", "noiseless", "rows", "-"]
dur = tl[:, :, 1] - tl[:, :, 0]
print(f"workload {name}: {ms / iters * 1e3:.1f} us per iteration over {iters} iterations (events){' — isolated, L2 flushed' if flush else ''}")
t0 = tl[:, 0, 0]
print("  start offsets from the sampler's first CTA (median us): " + "  ".join(
    f"{names[k]} {np.median(tl[:, k, 0] - t0):.1f}..{np.median(tl[:, k, 1] - t0):.1f}" for k in range(8) if not np.all(tl[:, k, 0] < 0)))
for k, n in enumerate(names):
    if np.all(tl[:, k, 0] < 0):
        continue
    print(f"  {n:10s} median {np.median(dur[:, k]):8.1f} us   min {dur[:, k].min():8.1f}   max {dur[:, k].max():8.1f}")
fused = bool(np.all(tl[:, 4, 0] < 0))          # single GPU: the apply step rides on the update kernel's last CTA
last = 3 if fused else 4
order = [k for k in [0, 1, 2, 3] + ([] if fused else [4]) if not np.all(tl[:, k, 0] < 0)]   # weights is absent when K7-K9 run as one kernel
gaps = [tl[:, b, 0] - tl[:, a, 1] for a, b in zip(order[:-1], order[1:])]
print("  gaps " + "->".join(names[k] for k in order) + " (median us):", [round(float(np.median(g)), 1) for g in gaps],
      "(cost = state kernel; the control rows ('reuse' slot) run beside it on a second stream)")
nxt = tl[1:, 0, 0] - tl[:-1, last, 1]
print(f"  gap {names[last]} -> next sample (median us):", round(float(np.median(nxt)), 1))
print("  iteration period (median us):", round(float(np.median(tl[1:, 0, 0] - tl[:-1, 0, 0])), 1))
if not np.all(tl[:, 5, 0] < 0):      # absent when the noise-less rollout rides on the state kernel (specialised kernel)
  print(f"  noiseless start after {names[last]} end (median us):", round(float(np.median(tl[:, 5, 0] - tl[:, last, 1])), 1),
      " noiseless end before the next weights / update start:", round(float(np.median(tl[1:, 2 if 2 in order else 3, 0] - tl[:-1, 5, 1])), 1))
