"""Multi-GPU plumbing: one process per GPU (torchrun), NCCL over NVLink for the single exchange step.

* Encoder: graphs are independent units -> contiguous shards per rank balanced by an edge-work estimate, NO collective
  (SURVEY.md section 8e).
* Train step: data parallel, every rank steps on its own batch; one sum all-reduce of the flat gradient buffer per
  step, rescaled by 1/world inside the Adam kernel (FlatAdam.all_reduce_grads / step_device).
The helpers below are backend-agnostic (the CPU test-suite drives them over gloo with world_size 2).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(costs, world):
    """Split range(len(costs)) into `world` contiguous shards with near-equal total cost. Returns world+1 offsets."""
    costs = np.asarray(costs, dtype=np.float64)
    total = costs.sum()
    if len(costs) == 0 or total <= 0:
        return np.linspace(0, len(costs), world + 1).round().astype(np.int64)
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    cuts = np.searchsorted(cum, total * np.arange(1, world) / world, side='left')
    return np.concatenate([[0], cuts, [len(costs)]]).astype(np.int64)


def encoder_cost(edge_ptr, node_ptr, h):
    """Per-graph work estimate of the encoder: edges x (nodes touched per edge ~ min(n, (avg degree)^h))."""
    e = np.diff(np.asarray(edge_ptr, dtype=np.float64))
    n = np.maximum(np.diff(np.asarray(node_ptr, dtype=np.float64)), 1.0)
    ball = np.minimum(n, np.maximum(e / n, 1.0) ** h + 1.0)
    return e * ball + n * n / 32.0


def my_shard(edge_ptr, node_ptr, h, world=None, rank=None):
    """(first_graph, last_graph_exclusive) of this rank."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
    b = shard_bounds(encoder_cost(edge_ptr, node_ptr, h), world)
    return int(b[rank]), int(b[rank + 1])


def allreduce_mean_(flat, group=None):
    """In-place mean of a flat gradient buffer over the group (sum all-reduce + scale)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / dist.get_world_size(group))
    return flat
