"""Graph-sharded data parallelism: one process per GPU, graphs of a batch split across ranks, ONE
all-reduce per step over a flat fp32 gradient buffer (NCCL over NVLink 5 / NVSwitch; gloo on CPU for
the host-logic tests).

The reference has no distributed path (single process, ``--devid``).  Graphs in a batch
are disjoint components (Batch.from_data_list, dataloader.py:11), so no message crosses shards and
the only exchange is the gradient sum.  The reference loss is a MEAN over all nodes of the batch
(train_botnet.py:225,287), so each rank back-propagates the SUM of its per-node losses and the
reduced gradient is divided by the GLOBAL node count: the step equals the single-GPU batched step.
"""
import torch
import torch.distributed as dist


def shard_range(num_items, world_size, rank):
    """contiguous balanced split: the first (num_items % world_size) ranks get one extra item"""
    base, extra = divmod(int(num_items), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_by_weight(weights, world_size):
    """Greedy longest-processing-time assignment of graphs (weights = edge counts) to ranks;
    returns a list of index lists, deterministic for equal inputs."""
    order = sorted(range(len(weights)), key=lambda i: (-int(weights[i]), i))
    loads = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += int(weights[i])
    return [sorted(v) for v in out]


class FlatGradientReducer:
    """Parameters' .grad tensors are views into one flat fp32 buffer with two extra slots
    (loss sum, sample count), so a step needs exactly one collective."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel + 2, dtype=torch.float32, device=dev)
        self.group = group
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("FlatGradientReducer expects fp32 parameters")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    @property
    def grads(self):
        return self.flat[:self.numel]

    def zero(self):
        self.flat.zero_()

    def reduce_mean(self, loss_sum, count):
        """All-reduce (sum) gradients, loss sum and sample count; scale gradients by 1/global_count.
        Returns (global mean loss, global count) as 0-dim tensors on the device (no host sync)."""
        self.flat[self.numel] = loss_sum.detach().to(torch.float32)
        self.flat[self.numel + 1] = float(count) if not torch.is_tensor(count) else count.to(torch.float32)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        total = self.flat[self.numel + 1].clone()
        mean_loss = self.flat[self.numel] / total
        self.flat[:self.numel].div_(total)
        return mean_loss, total


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*); returns
    (rank, local_rank, world_size).  A single process needs no process group."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off (Linux sysfs), so that the pinned host
    buffers it allocates afterwards (first touch) and its loader thread sit next to the GPU's PCIe root.  With 8 ranks
    copying a batch per step each, buffers that all live on one socket share that socket's memory controllers and the
    inter-socket link.  Returns the node id, or None when the topology cannot be read (single node, no sysfs, not
    Linux) — then nothing is changed."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(device_index), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        if bus is None:
            return None
        bdf = f"{dom:04x}:{bus:02x}:{dev:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
