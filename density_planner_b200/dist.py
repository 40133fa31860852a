"""Multi-GPU plumbing: one process per GPU, samples sharded in contiguous ranges, no collective in the rollout.

The reference has no distributed code at all; samples are independent given the shared reference
trajectory and weights (systems/ControlAffineSystem.py:441-450 has no cross-sample op), so the path shards naturally.
The only exchange steps are tiny next to the rollout:
  * density normaliser (simulation_objects.py:295-299): per time slot MAX of the log-density, then SUM of
    exp(l - max) * rho0 -> two allreduces of Ts values;
  * occupancy grids: SUM allreduce of the int64 fixed-point mass and int32 count grids -> bit-identical results at
    1/2/4/8 GPUs;
  * collision cost: SUM allreduce of the per-slice partial costs (fp64).
The merge logic is plain torch on whatever device the tensors live on, so it is covered by world_size-2 gloo tests on
the CPU; on the GPU box the backend is NCCL over NVLink.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, world, local_rank


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and hence its pinned host buffers, first-touch) to the CPUs of the NUMA node the GPU hangs off.

    With one process per GPU the host<->device copies of all ranks run at the same time; a rank whose pinned buffers sit
    on the other socket pays the inter-socket link for every byte.  Best effort: returns the node or None."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev_id = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev_id)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def shard_range(n_total, rank, world):
    """Contiguous, balanced sample range [lo, hi) of `rank`."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_normalizer_stats(lmax, lsum):
    """Per-slot (max, sum at that max) of the local shard -> global (max, sum at the global max).

    lmax fp32 [Ts], lsum fp64 [Ts]; returns new tensors.  With one process this is the identity."""
    if world_size() == 1:
        return lmax, lsum
    gmax = lmax.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
    scaled = lsum * torch.exp((lmax - gmax).to(torch.float64))
    # a shard whose max is -inf (no samples) contributes nothing
    scaled = torch.where(torch.isfinite(lmax), scaled, torch.zeros_like(scaled))
    dist.all_reduce(scaled, op=dist.ReduceOp.SUM)
    return gmax, scaled


def allreduce_grids(mass, count):
    """In-place SUM of the integer occupancy grids over all ranks (exact, order independent)."""
    if world_size() > 1:
        dist.all_reduce(mass, op=dist.ReduceOp.SUM)
        dist.all_reduce(count, op=dist.ReduceOp.SUM)
    return mass, count


def allreduce_sum(t):
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def max_over_ranks(value, device):
    """MAX of a python float over ranks (timing)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
