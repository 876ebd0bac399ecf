"""Graph batches and synthetic workloads — host side of the hot path.

``GraphBatch`` honours the batching contract of the reference (Batch.from_data_list,
src/gcn_meta/data/dataloader.py:11; kernel/train_eval.py:37-39): ``x``/``y`` concatenated along
dim 0, ``edge_index`` concatenated along dim 1 with a cumulative node offset, ``batch`` = graph id per
node (sorted ascending), ``slices_x`` = cumulative node boundaries (``batch.__slices__['x']``,
optim/train_eval_gc.py:17).  Synthetic generators are seeded numpy and deterministic, so the CPU
oracle and the GPU path see identical inputs.
"""
import numpy as np
import torch

BOTNET_NODES = 143107          # botnet_paper/botnet_plot.ipynb:408-409
BOTNET_EDGE_ENTRIES = 1_500_000
BOTNET_EVIL = 10_000


class GraphBatch:
    """x [N,F] f32, edge_index [2,E] i64, y, batch [N] i64, slices_x [G+1] (host list)."""

    def __init__(self, x, edge_index, y=None, batch=None, slices_x=None, slices_e=None):
        self.x, self.edge_index, self.y = x, edge_index, y
        if batch is None:
            if slices_x is None:
                slices_x = [0, int(x.size(0))]
        elif slices_x is None:  # boundaries from the (sorted) graph-id vector
            counts = torch.bincount(batch.cpu()) if batch.numel() else torch.zeros(0, dtype=torch.long)
            slices_x = [0] + torch.cumsum(counts, 0).tolist()
        self._batch = batch     # None: built from the boundaries when first asked for (node-level models never do)
        self.slices_x = list(slices_x)
        self.slices_e = list(slices_e) if slices_e is not None else [0, int(edge_index.size(1))]
        self.__slices__ = {"x": self.slices_x, "edge_index": self.slices_e}

    @property
    def batch(self):
        if self._batch is None:
            counts = torch.tensor([b - a for a, b in zip(self.slices_x, self.slices_x[1:])], dtype=torch.long,
                                  device=self.x.device)
            self._batch = torch.repeat_interleave(torch.arange(counts.numel(), device=self.x.device), counts)
        return self._batch

    @batch.setter
    def batch(self, value):
        self._batch = value

    @property
    def num_graphs(self):
        return len(self.slices_x) - 1

    @property
    def num_nodes(self):
        return int(self.x.size(0))

    @property
    def num_edges(self):
        return int(self.edge_index.size(1))

    @property
    def num_features(self):
        return int(self.x.size(1))

    def _map(self, fn):
        out = GraphBatch(fn(self.x), fn(self.edge_index), None if self.y is None else fn(self.y),
                         None if self._batch is None else fn(self._batch), self.slices_x, self.slices_e)
        return out

    def to(self, device, non_blocking=False):
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def structure(self, loop_mode=0, recycle=False):
        """the row structures of this batch's edge_index (meta_gcn_b200.graph.GraphStructure), registered in the
        structure cache together with the batch boundaries, so that the model classes — which only see edge_index —
        find it: an ordered batch (every graph in the reference's preprocessed layout) then needs no sort"""
        from .graph import structure_of
        return structure_of(self.edge_index, self.num_nodes, loop_mode, segments=(self.slices_x, self.slices_e),
                            recycle=recycle)

    def with_int32_indices(self):
        """edge_index as int32 (N, E < 2^31): half the bytes of the host -> device copy, which is what bounds a
        botnet step end to end (600 MB of int64 indices per 25-graph batch, train_botnet.py:282).  Convert once when
        the dataset is loaded; every structure-building entry point accepts either dtype."""
        if self.edge_index.dtype == torch.int32:
            return self
        if self.num_nodes >= 2 ** 31 - 2 ** 20:
            raise ValueError("too many nodes for int32 indices")
        return GraphBatch(self.x, self.edge_index.to(torch.int32), self.y, self._batch, self.slices_x, self.slices_e)

    @staticmethod
    def from_data_list(graphs):
        """graphs: list of dicts / objects with x, edge_index, y (numpy arrays or tensors, on the host or all on one
        CUDA device).  Batch.from_data_list of data/dataloader.py:11: node-wise concatenation, edge indices shifted by
        the node offset, graph-id vector, cumulative boundaries.  Device inputs are collated on the device (one
        cat per field, offsets added by a broadcast) — no host round trip."""
        xs, eis, ys, sx, se = [], [], [], [0], [0]
        for item in graphs:
            get = item.get if isinstance(item, dict) else (lambda k, it=item: getattr(it, k, None))
            x = torch.as_tensor(get("x"))
            ei = torch.as_tensor(get("edge_index")).long()
            xs.append(x)
            eis.append(ei)
            y = get("y")
            if y is not None:
                ys.append(torch.as_tensor(y).reshape(-1) if np.ndim(y) <= 1 else torch.as_tensor(y))
            sx.append(sx[-1] + x.size(0))
            se.append(se[-1] + ei.size(1))
        dev = xs[0].device if xs else torch.device("cpu")
        counts_n = torch.tensor([b - a for a, b in zip(sx, sx[1:])], dtype=torch.long, device=dev)
        counts_e = torch.tensor([b - a for a, b in zip(se, se[1:])], dtype=torch.long, device=dev)
        gid = torch.arange(len(xs), dtype=torch.long, device=dev)
        batch = torch.repeat_interleave(gid, counts_n)                                   # graph id per node, sorted
        node_off = torch.tensor(sx[:-1], dtype=torch.long, device=dev)
        edge_off = torch.repeat_interleave(node_off, counts_e)                           # node offset per edge entry
        edge_index = (torch.cat(eis, dim=1) + edge_off.unsqueeze(0)) if eis else torch.zeros(2, 0, dtype=torch.long)
        y = torch.cat(ys) if ys else None
        x = torch.cat(xs) if xs else torch.zeros(0, 0)
        return GraphBatch(x, edge_index, y, batch, sx, se)


class DeviceLoader:
    """Prefetching loader over pinned host GraphBatches (the role of num_workers in
    src/gcn_meta/data/dataloader.py:6-30 and of `batch.to(device)` at train_botnet.py:282): the H2D copy of
    batch k+1 runs on a second stream while batch k computes, into `slots` preallocated device slots per field,
    so the steady state allocates nothing (a fresh 600 MB allocation per step costs a cudaMalloc and its
    device synchronisation).  A slot is overwritten only after the work enqueued for its previous batch.

        for batch in DeviceLoader(host_batches, "cuda"):      # host_batches: iterable of GraphBatch
            loss = step(batch)
    """

    FIELDS = ("x", "edge_index", "y", "batch")

    def __init__(self, batches, device, slots=2, fields=None):
        self.batches = batches
        if fields is not None:          # e.g. ("x", "edge_index", "y"): node-level models never read `batch`
            self.FIELDS = tuple(fields)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceLoader copies to a CUDA device (no CPU path)")
        self.slots = [dict() for _ in range(max(2, int(slots)))]
        self.stream = torch.cuda.Stream(device=self.device)

    def _copy(self, host, k):
        """enqueue host -> slot k on the copy stream; returns (device batch, event)"""
        slot = self.slots[k]
        cur = torch.cuda.current_stream(self.device)
        free = torch.cuda.Event()
        free.record(cur)                       # everything enqueued so far (incl. this slot's previous batch)
        self.stream.wait_event(free)
        out = {}
        with torch.cuda.stream(self.stream):
            for name in self.FIELDS:
                src = getattr(host, name)
                if src is None:
                    out[name] = None
                    continue
                dst = slot.get(name)
                if dst is None or dst.shape != src.shape or dst.dtype != src.dtype:
                    dst = torch.empty(src.shape, dtype=src.dtype, device=self.device)
                    slot[name] = dst
                dst.copy_(src, non_blocking=True)
                out[name] = dst
            ev = torch.cuda.Event()
            ev.record(self.stream)
        b = GraphBatch(out["x"], out["edge_index"], out.get("y"), out.get("batch"), host.slices_x, host.slices_e)
        # the (lazy) structure object carries the batch boundaries; nothing is launched here.  It takes over the
        # buffers of the structure this slot's previous batch had (same shapes): no allocation in the steady state,
        # and the same addresses, which is what lets a CUDA graph captured for the slot be replayed
        b.structure(recycle=True)
        return b, ev

    def bytes_per_batch(self, host):
        """bytes one batch moves host -> device (what the copies above transfer)"""
        return sum(t.numel() * t.element_size() for t in (getattr(host, n) for n in self.FIELDS) if t is not None)

    def __iter__(self):
        it = iter(self.batches)
        try:
            pending = self._copy(next(it), 0)
        except StopIteration:
            return
        k = 0
        while pending is not None:
            b, ev = pending
            torch.cuda.current_stream(self.device).wait_event(ev)
            k += 1
            try:
                pending = self._copy(next(it), k % len(self.slots))
            except StopIteration:
                pending = None
            yield b


# ------------------------------------------------------------------------------------------------
# edge preprocessing with the reference's ordering (data_procs/undirected.py:6-35, loop.py:13-17,
# data_add_degree.py:45-65): symmetrise, sort-unique by (src,dst), append loops, out-degree
# ------------------------------------------------------------------------------------------------
def symmetrise_sorted_with_loops(src, dst, num_nodes):
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    key = np.concatenate([src * num_nodes + dst, dst * num_nodes + src])
    key = np.unique(key)  # sorted: lexicographic (src, dst)
    loop = np.arange(num_nodes, dtype=np.int64)
    row = np.concatenate([key // num_nodes, loop])
    col = np.concatenate([key % num_nodes, loop])
    return np.stack([row, col])


def synth_botnet_graph(seed=0, num_nodes=BOTNET_NODES, edge_entries=BOTNET_EDGE_ENTRIES,
                       evil=BOTNET_EVIL):
    """C1: power-law background traffic graph + planted ring-with-chords P2P botnet (y=1),
    symmetrised, sort-unique, self-loops appended; x = [1, out-degree] as the botnet HDF5 stores it
    (botnet_plot.ipynb:382: x = [[1.,313.],[1.,3.],...])."""
    rng = np.random.default_rng(seed)
    n = int(num_nodes)
    evil = min(int(evil), n // 2)
    # botnet overlay: ring + power-of-two chords (Chord-like fingers)
    bots = rng.choice(n, size=evil, replace=False)
    bs, bd = [], []
    if evil >= 3:
        idx = np.arange(evil)
        for hop in (1, 2, 4, 8, 64, 512):
            if hop < evil:
                bs.append(bots[idx])
                bd.append(bots[(idx + hop) % evil])
    bs = np.concatenate(bs) if bs else np.zeros(0, np.int64)
    bd = np.concatenate(bd) if bd else np.zeros(0, np.int64)
    undirected_target = max((int(edge_entries) - n) // 2 - len(bs), 0)
    # background: one endpoint by power-law popularity, partner uniform
    pop = (np.arange(1, n + 1, dtype=np.float64)) ** -0.8
    pop /= pop.sum()
    ids = rng.permutation(n)
    m = int(undirected_target * 1.035) + 16  # head-room for duplicates / loops removed below
    a = ids[rng.choice(n, size=m, p=pop)]
    b = rng.integers(0, n, size=m)
    ok = a != b
    a, b = a[ok], b[ok]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    _, first = np.unique(lo * n + hi, return_index=True)
    first = np.sort(first)[:undirected_target]
    a, b = a[first], b[first]
    ei = symmetrise_sorted_with_loops(np.concatenate([a, bs]), np.concatenate([b, bd]), n)
    deg = np.bincount(ei[0], minlength=n).astype(np.float32)
    x = np.stack([np.ones(n, dtype=np.float32), deg], axis=1)
    y = np.zeros(n, dtype=np.uint8)
    y[bots] = 1
    return {"x": x, "edge_index": ei, "y": y}


def synth_tu_graph(rng, f_in=3, num_classes=2):
    """C3: PROTEINS/ENZYMES-like small graph: ~40 nodes (clipped normal), undirected avg degree
    3.7, random-normal features (kernel/plot.ipynb:45,245 show x=[45,3]/[20,3])."""
    n = int(np.clip(round(rng.normal(40, 25)), 4, 620))
    m = max(int(round(n * 3.7 / 2)), 1)
    a = rng.integers(0, n, size=m)
    b = rng.integers(0, n, size=m)
    keep = a != b
    a, b = a[keep], b[keep]
    key = np.unique(np.concatenate([a * n + b, b * n + a]))
    ei = np.stack([key // n, key % n]).astype(np.int64)
    x = rng.normal(size=(n, f_in)).astype(np.float32)
    return {"x": x, "edge_index": ei, "y": np.array([rng.integers(0, num_classes)], dtype=np.int64)}


def synth_tu_batch(seed=0, num_graphs=128, f_in=3, num_classes=2):
    rng = np.random.default_rng(seed)
    return GraphBatch.from_data_list([synth_tu_graph(rng, f_in, num_classes) for _ in range(num_graphs)])


def synth_powerlaw_graph(seed, num_nodes, num_edges, alpha=0.8, symmetric=False):
    """C4/C5: directed power-law edge list (popular sources, uniform targets), unsorted order."""
    rng = np.random.default_rng(seed)
    n = int(num_nodes)
    pop = (np.arange(1, n + 1, dtype=np.float64)) ** -alpha
    pop /= pop.sum()
    cdf = np.cumsum(pop)
    ids = rng.permutation(n)
    src = ids[np.minimum(np.searchsorted(cdf, rng.random(int(num_edges))), n - 1)]
    dst = rng.integers(0, n, size=int(num_edges))
    if symmetric:
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
    return np.stack([src, dst]).astype(np.int64)


class _Meta:
    """dataset stand-in for the kernel/ nets' constructor (dataset.num_features / num_classes)"""

    def __init__(self, num_features, num_classes):
        self.num_features = num_features
        self.num_classes = num_classes


def dataset_meta(num_features, num_classes):
    return _Meta(num_features, num_classes)
