"""Multi-GPU for the rollout path: env shards per rank, ONE gradient all-reduce per optimizer step.

The reference has no distributed code.  Envs are independent, so rank g owns envs
[g*N/G, (g+1)*N/G) seeded by GLOBAL env index (``env_index_offset``); rollout, GAE, advantage normalisation and
minibatching are rank-local (identical to running the reference per shard); the policy is replicated and its
gradients are averaged with a single NCCL all-reduce over one flat fp32 bucket (NVLink 5 / NVSwitch; the bucket
is tens of KB for the MLP, 6.7 MB for NatureCNN -- latency-bound, so exactly one collective, no bucketing).
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


def shard(num_envs_total, rank, world):
    """Contiguous env shard of this rank: (env_index_offset, local num_envs)."""
    if num_envs_total % world != 0:
        raise ValueError('total num_envs must be divisible by the number of ranks')
    per = num_envs_total // world
    return rank * per, per


class GradBucket:
    """Flat fp32 gradient bucket: parameters' ``.grad`` are views into one buffer, so the all-reduce needs no
    pack/unpack copies."""

    def __init__(self, module):
        params = [p for p in module.parameters() if p.requires_grad]
        total = sum(p.numel() for p in params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
        off = 0
        for p in params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        self.params = params
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    def zero(self):
        self.rebind()
        self.flat.zero_()

    def rebind(self):
        """optimizer.zero_grad(set_to_none=True) drops the views: re-attach them (cheap, no copies)."""
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * 4:
                g = self.flat[off:off + n].view_as(p)
                if p.grad is not None:
                    g.copy_(p.grad)
                p.grad = g
            off += n

    def all_reduce_mean(self):
        self.rebind()
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(self.world)


def broadcast_parameters(module, src=0):
    if dist.is_initialized() and dist.get_world_size() > 1:
        for p in module.parameters():
            dist.broadcast(p.data, src)
        for b in module.buffers():
            dist.broadcast(b.data, src)
