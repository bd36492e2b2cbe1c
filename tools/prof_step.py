import sys, os, torch, time
sys.path.insert(0, '/root/repo')
import bench
import transformer_transducer_b200 as ttb
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda:0")
torch.manual_seed(1234)
joint = (ttb.JointNet(2 * w["D"], w["H"], w["V"]) if w["joint"] == "tt" else
         ttb.JointNetwork(w["V"], w["D"], w["D"], w["H"], "tanh")).to(dev)
crit = ttb.RNNTLoss(blank=0, reduction="mean")
enc, pred, labels, al, ll = bench.synth(w, 1234, device=dev)
enc.requires_grad_(); pred.requires_grad_()
def step():
    for p_ in joint.parameters(): p_.grad = None
    enc.grad = None; pred.grad = None
    loss = crit(joint(enc, pred) if w["joint"] == "tt" else joint(enc[:, :, None], pred[:, None]), labels, al, ll)
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
# gaps on the GPU timeline of the last step
t_end = 0; gaps = 0; busy = 0
first = ev[0].time_range.start
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    if s > t_end and t_end > 0: gaps += s - t_end
    busy += t - s
    t_end = max(t_end, t)
print("GPU span us", t_end - first, "busy", busy, "gaps", gaps, "per step gaps", gaps / 3)
# the largest idle intervals of the timeline (all streams merged), with the kernels on either side
iv = []
t_end, prev = 0, None
for e in ev:
    s_, t_ = e.time_range.start, e.time_range.end
    if t_end and s_ > t_end:
        iv.append((s_ - t_end, prev, e.name))
    if t_ > t_end:
        t_end, prev = t_, e.name
for gap, a, b in sorted(iv, reverse=True)[:24]:
    print("%8.1f us   %-50s -> %s" % (gap, a[:50], b[:60]))
