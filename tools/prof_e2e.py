"""GPU-timeline gaps of the END-TO-END step (fresh device tensors every step: H2D prefetch on a copy stream, the length
check's host read, loss.item()), configs[1].  Prints the idle intervals of the merged timeline; run under gpurun."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import transformer_transducer_b200 as ttb
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda:0")
torch.manual_seed(1234)
joint = ttb.JointNetwork(w["V"], w["D"], w["D"], w["H"], "tanh").to(dev)
crit = ttb.RNNTLoss(blank=0, reduction="mean")
host = bench.synth(w, 1234, pin=True)
copy_stream = torch.cuda.Stream(dev)
def fetch():
    with torch.cuda.stream(copy_stream):
        ts = [t.to(dev, non_blocking=True) for t in host]
        ev = torch.cuda.Event(); ev.record(copy_stream)
    return ts, ev
pending = []
def step():
    for p_ in joint.parameters(): p_.grad = None
    (e, p_, lab, al, ll), ev = pending.pop() if pending else fetch()
    torch.cuda.current_stream(dev).wait_event(ev)
    e.requires_grad_(); p_.requires_grad_()
    loss = crit(joint(e[:, :, None], p_[:, None]), lab, al, ll)
    pending.append(fetch())
    loss.backward()
    return loss.item()
for _ in range(4): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
N = 4
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(N): step()
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy HtoD" not in e.name],
            key=lambda e: e.time_range.start)
iv, t_end, prev, first = [], 0, None, ev[0].time_range.start
for e in ev:
    s_, t_ = e.time_range.start, e.time_range.end
    if t_end and s_ > t_end: iv.append((s_ - t_end, prev, e.name))
    if t_ > t_end: t_end, prev = t_, e.name
print("span per step %.1f us, idle per step %.1f us (compute streams only, H2D copies excluded)" % ((t_end - first) / N, sum(g for g, _, _ in iv) / N))
for gap, a, b in sorted(iv, reverse=True)[:16]:
    print("%8.1f us   %-48s -> %s" % (gap, a[:48], b[:56]))
