"""One process per GPU: env instances shard by contiguous global-id blocks; the only collective is the
sum all-reduce of the episode-statistics vector.  Uses torch.distributed (NCCL on
GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as td


def shard_range(total_envs, rank, world_size):
    """Contiguous block of global env ids owned by `rank`: [lo, hi).  Blocks differ by at most one env."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def bind_to_gpu_cpus(device_index):
    """Restricts this process to the host cores next to its GPU (NVML's CPU affinity of the device), so that the pinned
    host buffers it allocates -- the targets of the D2H copies of the host-buffer path -- and its host threads live on the
    GPU's NUMA node.  With eight processes on a two-socket box half of the copies cross the socket link otherwise.
    Returns the core list, or None when NVML has nothing to say (virtual machines often report every core) or
    CS_NO_CPU_BIND is set."""
    if os.environ.get("CS_NO_CPU_BIND"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed or len(allowed) >= len(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def init_from_env(backend=None, bind_cpus=False):
    """Initialises torch.distributed from RANK/WORLD_SIZE/MASTER_* when launched by torchrun.  bind_cpus: also restrict
    the process to the cores next to its GPU (bind_to_gpu_cpus) when there are several ranks."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if bind_cpus and world > 1 and torch.cuda.is_available():
        bind_to_gpu_cpus(local)
    if world > 1 and not td.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        if backend == "nccl":
            torch.cuda.set_device(local)
            td.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            td.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def allreduce_stats(stats, group=None):
    """Sum of the per-GPU statistics vectors ([CS_NUM_STATS] float64).  Accepts the device tensor
    (``env.stats_tensor``) or anything convertible; returns a new tensor, the input is untouched."""
    t = torch.as_tensor(stats, dtype=torch.float64).clone()
    if td.is_available() and td.is_initialized() and td.get_world_size(group) > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM, group=group)
    return t


def max_over_ranks(value, device=None):
    """Max of a python float over ranks (timing: the slowest rank defines the step time)."""
    if not (td.is_available() and td.is_initialized()) or td.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())
