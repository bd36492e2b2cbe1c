"""GPU timeline of the DDP step (rank 0): python -m torch.distributed.run --nproc-per-node N tools/prof_ddp.py [cfg2]
Prints the idle intervals of the merged GPU timeline with the kernels on either side, and the NCCL kernels' placement."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import transformer_transducer_b200 as ttb
import torch.distributed as dist
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(1234)
joint = ttb.JointNetwork(w["V"], w["D"], w["D"], w["H"], "tanh").to(dev)
model = torch.nn.parallel.DistributedDataParallel(joint, device_ids=[local], gradient_as_bucket_view=True, bucket_cap_mb=8)
crit = ttb.RNNTLoss(blank=0, reduction="mean")
enc, pred, labels, al, ll = bench.synth(w, 1234 + rank, device=dev)
enc.requires_grad_(); pred.requires_grad_()
def step():
    for p_ in joint.parameters(): p_.grad = None
    enc.grad = None; pred.grad = None
    loss = crit(model(enc[:, :, None], pred[:, None]), labels, al, ll)
    loss.backward()
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
if rank == 0: print("unprofiled: %.3f ms / step" % (e0.elapsed_time(e1) / 10))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
if rank == 0:
    ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    first, t_end, prev, iv, busy = ev[0].time_range.start, 0, None, [], 0
    for e in ev:
        s_, t_ = e.time_range.start, e.time_range.end
        if t_end and s_ > t_end: iv.append((s_ - t_end, prev, e.name))
        if t_ > t_end: t_end, prev = t_, e.name
    print("span %.1f us for 3 steps, idle %.1f us per step" % (t_end - first, sum(g for g, _, _ in iv) / 3))
    for gap, a, b in sorted(iv, reverse=True)[:16]:
        print("%8.1f us   %-46s -> %s" % (gap, a[:46], b[:60]))
    for e in ev:
        if "nccl" in e.name.lower():
            print("nccl %-40s start %+9.1f us  dur %7.1f" % (e.name[:40], e.time_range.start - first, e.time_range.end - e.time_range.start))
    cpu = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.time_range.end - e.time_range.start > 300],
                 key=lambda e: -(e.time_range.end - e.time_range.start))[:14]
    for e in cpu: print("cpu %-60s %8.1f us" % (e.name[:60], e.time_range.end - e.time_range.start))
dist.destroy_process_group()
