"""world_size-2 gloo test of the N>1 path: the lazy logits handle crosses DistributedDataParallel's forward
untouched and the joint's parameter gradients come back averaged over ranks (the only collective on the path)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import transformer_transducer_b200 as ttb
    from oracle import rnnt_oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)

    class LazyOnCpu(ttb.JointNetwork):  # same split as the CUDA branch of JointNetwork.forward, forced on CPU
        def forward(self, h_enc, h_dec):
            return ttb.LazyJointLogits(self.lin_enc(h_enc.squeeze(2)), self.lin_dec(h_dec.squeeze(1)),
                                       self.lin_out.weight, self.lin_out.bias)

    joint = LazyOnCpu(11, 8, 8, 16, "tanh")
    ddp = torch.nn.parallel.DistributedDataParallel(joint)
    g = torch.Generator().manual_seed(100 + rank)      # each rank owns a different shard of the batch
    enc, pred = torch.randn(2, 6, 8, generator=g), torch.randn(2, 4, 8, generator=g)
    labels = torch.randint(1, 11, (2, 3), generator=g, dtype=torch.int32)
    al, ll = torch.tensor([6, 5], dtype=torch.int32), torch.tensor([3, 2], dtype=torch.int32)
    z = ddp(enc[:, :, None], pred[:, None])
    assert isinstance(z, ttb.LazyJointLogits)
    loss = rnnt_oracle.RNNTLoss()(z.materialize(), labels, al, ll)   # CPU stand-in for the CUDA loss
    loss.backward()
    grads = torch.cat([p.grad.reshape(-1) for p in joint.parameters()])
    gathered = [torch.zeros_like(grads) for _ in range(world)]
    dist.all_gather(gathered, grads)
    same = all(torch.allclose(gathered[0], x) for x in gathered)
    # reference: single process over the concatenated global batch with 'mean' == average of per-rank means
    ret[rank] = (same, grads.clone(), loss.detach().clone(), (enc, pred, labels, al, ll))
    dist.destroy_process_group()


def test_ddp_gloo_world2():
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0][0] and ret[1][0]
    sys.path.insert(0, ROOT)
    import transformer_transducer_b200 as ttb
    from oracle import rnnt_oracle
    torch.manual_seed(0)
    joint = ttb.JointNetwork(11, 8, 8, 16, "tanh")
    total = 0
    for r in range(2):
        enc, pred, labels, al, ll = ret[r][3]
        total = total + rnnt_oracle.RNNTLoss()(joint(enc[:, :, None], pred[:, None]), labels, al, ll) / 2
    total.backward()
    want = torch.cat([p.grad.reshape(-1) for p in joint.parameters()])
    assert torch.allclose(ret[0][1], want, atol=1e-5), float((ret[0][1] - want).abs().max())
