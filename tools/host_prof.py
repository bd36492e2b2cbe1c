"""Host-side cost of one joint + loss step: issue time per step (no synchronisation) and a cProfile of the Python front end.
    python tools/host_prof.py [cfg1|cfg2|...]      (run under gpurun)"""
import os, sys, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import transformer_transducer_b200 as ttb
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg1"]
dev = torch.device("cuda:0")
torch.manual_seed(1234)
tt = w["joint"] == "tt"
joint = (ttb.JointNet(2 * w["D"], w["H"], w["V"]) if tt else ttb.JointNetwork(w["V"], w["D"], w["D"], w["H"], "tanh")).to(dev)
crit = ttb.RNNTLoss(blank=0, reduction="mean")
enc, pred, labels, al, ll = bench.synth(w, 1234, device=dev)
enc.requires_grad_(); pred.requires_grad_()
def step():
    for p_ in joint.parameters(): p_.grad = None
    enc.grad = None; pred.grad = None
    loss = crit(joint(enc, pred) if tt else joint(enc[:, :, None], pred[:, None]), labels, al, ll)
    loss.backward()
for _ in range(5): step()
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for _ in range(N): step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("host issue time per step %.1f us, wall per step %.1f us" % (1e6 * t_host / N, 1e6 * t_all / N))
pr = cProfile.Profile(); pr.enable()
for _ in range(N): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
