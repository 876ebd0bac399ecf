"""CUDA-graph replay of a training step's device work for batches of a FIXED shape.

The botnet step is ~120 kernel launches for 23 ms of device time, so the host normally stays ahead — but any host
jitter (a busy box, the Python GC) shows up as idle gaps, and a single graph (C1, 2.3 ms per step) is plainly
launch-bound.  Every entry point of libmgcn.so is capture-safe by contract (no allocation, no synchronisation, no
host reads: include/mgcn.h), so forward + loss + backward can be recorded once and replayed:

    g = GraphedCall(lambda: fwd_loss_bwd(batch))     # zero grads, forward, loss, backward; returns the loss tensor
    loss = g()                                       # replay; `loss` is the captured output tensor (overwritten per replay)

Inputs are the tensors the closure reads (static addresses: copy new data INTO them); parameter gradients must
already exist (they are accumulated in place).  Shapes, the edge structure and the set of launches are frozen at
capture time — a batch of another shape needs its own GraphedCall.
"""
import torch


class GraphedCall:
    """Run the training loop (eager warm-up steps included) under ONE non-default stream and construct this object
    there: autograd's AccumulateGrad nodes remember the stream they were created on, and a node created on the legacy
    default stream cannot run inside a capture (cudaErrorStreamCaptureImplicit).  Called on the default stream, a
    private side stream is used for warm-up and capture, which works as long as no earlier eager backward on the
    default stream is still referenced."""

    def __init__(self, fn, warmup=3):
        self.fn = fn
        self.graph = None
        self.out = None
        cur = torch.cuda.current_stream()
        if cur == torch.cuda.default_stream():
            side = torch.cuda.Stream()
            side.wait_stream(cur)
        else:
            side = cur
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):       # lazy initialisation (cudaFuncSetAttribute, structure cache, allocator)
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            self.out = fn()
        self.graph = graph

    def __call__(self):
        self.graph.replay()
        return self.out


class GraphedSlots:
    """One captured graph per distinct set of input ADDRESSES — for batches that arrive in the preallocated slots of a
    `DeviceLoader` (meta_gcn_b200/data.py): a slot's tensors and (with recycled structure buffers) its row structures
    keep their addresses from batch to batch, so the graph captured for the slot the first time it is seen can be
    replayed for every later batch of the same shapes that lands in it.

        graphs = GraphedSlots(fwd_loss_bwd)               # fwd_loss_bwd(batch) -> loss tensor
        for batch in DeviceLoader(host_batches, "cuda"):
            gs = batch.structure(recycle=True); gs.fwd_plain; gs.bwd_plain     # built eagerly (one host read)
            loss = graphs(batch, extra_key=(...))        # capture on first sight of these addresses, replay after

    The key is made of the data pointers and shapes of the batch fields and of every built structure buffer; anything
    else the step depends on and that may change between batches (a branch taken on the data, say) goes into
    `extra_key`.  A key that was never seen costs an eager warm-up, a capture and a replay."""

    def __init__(self, fn, warmup=1, capacity=4):
        self.fn = fn
        self.warmup = warmup
        self.capacity = capacity
        self.graphs = {}

    @staticmethod
    def key_of(batch):
        parts = []
        for t in (batch.x, batch.edge_index, batch.y):
            if t is not None:
                parts.append((t.data_ptr(), tuple(t.shape), t.dtype))
        for csr in batch.structure().built():
            parts.append(tuple(t.data_ptr() for t in csr.tensors()))
        return tuple(parts)

    def __call__(self, batch, extra_key=()):
        key = (self.key_of(batch), extra_key)
        g = self.graphs.get(key)
        if g is None:
            while len(self.graphs) >= self.capacity:
                self.graphs.pop(next(iter(self.graphs)))
            g = GraphedCall(lambda: self.fn(batch), warmup=self.warmup)
            self.graphs[key] = g
        return g()
