"""Instance -> rank partitioning for data-parallel inference and training.

EV is block-diagonal per instance (instance_loader.py:56-66) and the parameters are shared,
so instances are independent units: each rank plans and runs its own sub-batch and the only
exchanges are one all-reduce of the zero-padded [B] logits vector per forward pass and one
all-reduce of the flat gradient blob (115,529 floats) per training step.
"""
import numpy as np


def partition_instances(n_edges, world_size):
    """Greedy longest-processing-time assignment balanced by edge count (work ~ sum of edges).

    Returns a list of ``world_size`` sorted index arrays covering range(len(n_edges)).
    """
    n_edges = np.asarray(n_edges, dtype=np.int64)
    order = np.argsort(-n_edges, kind="stable")
    loads = np.zeros(world_size, dtype=np.int64)
    parts = [[] for _ in range(world_size)]
    for k in order:
        r = int(np.argmin(loads))
        parts[r].append(int(k))
        loads[r] += n_edges[k]
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def take_instances(idx, EV_src, EV_dst, W, C, n_vertices, n_edges):
    """Extracts the sub-batch made of instances ``idx`` (kept in order) with local ids."""
    n_vertices = np.asarray(n_vertices, dtype=np.int64)
    n_edges = np.asarray(n_edges, dtype=np.int64)
    eoff = np.concatenate([[0], np.cumsum(n_edges)])
    voff = np.concatenate([[0], np.cumsum(n_vertices)])
    W = np.asarray(W).reshape(-1)
    C = np.asarray(C).reshape(-1)
    srcs, dsts, ws, cs = [], [], [], []
    vacc = 0
    for k in idx:
        e0, e1 = eoff[k], eoff[k + 1]
        shift = vacc - voff[k]
        srcs.append(EV_src[e0:e1].astype(np.int64) + shift)
        dsts.append(EV_dst[e0:e1].astype(np.int64) + shift)
        ws.append(W[e0:e1])
        cs.append(C[e0:e1])
        vacc += n_vertices[k]
    cat = lambda lst, dt: (np.concatenate(lst) if lst else np.zeros(0)).astype(dt)
    return (cat(srcs, np.int32), cat(dsts, np.int32), cat(ws, np.float32), cat(cs, np.float32),
            n_vertices[idx].astype(np.int32), n_edges[idx].astype(np.int32))


def scatter_logits(local_logits, idx, batch_size):
    """Zero-padded [B] vector holding this rank's logits at its instances' positions; the
    sum over ranks (all-reduce) is the full logits vector."""
    full = np.zeros(batch_size, dtype=np.float32)
    full[np.asarray(idx, dtype=np.int64)] = np.asarray(local_logits, dtype=np.float32)
    return full


def all_reduce_logits(local_logits, idx, batch_size, device=None):
    """One all-reduce(sum) over the default process group (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    full = torch.from_numpy(scatter_logits(local_logits, idx, batch_size))
    if device is not None:
        full = full.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM)
    return full


def all_reduce_gradients(grads, loss=None):
    """One all-reduce(sum), in place, of the flat gradient blob (and optionally the loss scalar)
    over the default process group.  Each rank computed its blob with the GLOBAL batch size as
    the divisor of the loss mean (tspgnn_backward's ``global_batch``), so the sum is the gradient
    of model.py:157's reduce_mean over the whole batch; the global-norm clip and Adam then run
    identically on every rank (tspgnn_apply_gradients), keeping the replicas bit-equal."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(grads, op=dist.ReduceOp.SUM)
        if loss is not None:
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    return grads, loss


def train_step_sharded(engine, dW, dC, d_route_exists, time_steps, global_batch):
    """Data-parallel training step of one rank: forward + reverse pass on this rank's instances,
    gradient all-reduce, optimizer.  Returns (loss of the whole batch, global gradient norm)."""
    import torch
    engine.train_forward(dW, dC, time_steps)
    loss, grads = engine.backward(d_route_exists, global_batch)
    # The collective must be ordered on the engine's stream: torch's NCCL group synchronises its own
    # stream with the CURRENT stream on both sides of the call, and the kernels that produce and
    # consume the blob run on engine.stream(), not on torch's default stream.
    with torch.cuda.stream(engine.stream()):
        all_reduce_gradients(grads, loss)
        loss_host = loss.cpu()
    gnorm = engine.apply_gradients(grads)          # synchronises the stream
    return float(loss_host[0]), gnorm
