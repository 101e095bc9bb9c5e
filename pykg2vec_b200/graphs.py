"""CUDA-graph staging of a host-in / host-out call.

The hot loop issues a handful of short kernels per batch (train step: 2 kernels + a
memset; evaluation batch: ~8), so launch latency and per-call host work dominate the
end-to-end time.  StagedGraph captures the whole round trip — ONE H2D copy from a pinned
staging buffer, the kernels, ONE D2H copy into a pinned result buffer — as a single CUDA
graph that is replayed per batch (Guideline: capture launch-bound inner loops in graphs).
The C-ABI entry points are capture-safe: they only enqueue work on the given stream.
"""
import torch


class StagedGraph:
    def __init__(self, device, in_words, out_like, body, in_dtype=torch.int64, warm_body=None, pre=None,
                 capture_error_mode="global"):
        """body(d_in) must enqueue the kernels on the current stream and return a device tensor
        shaped like `out_like` (a CPU tensor prototype) that it fully overwrites.  The staging
        buffer starts zero-filled (id 0 is always valid) and the un-captured warm-up run uses
        `warm_body` when the real body has side effects (a training step).
        pre(d_in): optional work that must stay OUTSIDE the graph — a collective on the staged inputs
        (data-parallel id exchange): then the H2D copy and pre() are issued eagerly by every call and
        the graph holds body + the D2H copy."""
        self.device = torch.device(device)
        self.h_in = torch.zeros(in_words, dtype=in_dtype).pin_memory()
        self.warm_body = warm_body
        self.d_in = torch.zeros(in_words, dtype=in_dtype, device=self.device)
        self.h_out = torch.empty_like(out_like).pin_memory()
        self.body = body
        self.pre = pre
        self.capture_error_mode = capture_error_mode
        self.graph = None
        self._done = None   # event of the last asynchronous replay

    def _stage(self):
        self.d_in.copy_(self.h_in, non_blocking=True)
        if self.pre is not None:
            self.pre(self.d_in)

    def _compute(self, body=None):
        d_out = (body or self.body)(self.d_in)
        self.h_out.copy_(d_out, non_blocking=True)

    def _run_eager(self, body=None):
        """the whole call without a graph (config.cuda_graph = False)"""
        self._stage()
        self._compute(body)

    def capture(self):
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):   # warm-up outside capture (loads kernels, sets attributes)
            self._stage()
            self._compute(self.warm_body)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode=self.capture_error_mode):
            if self.pre is None:
                self._stage()
            self._compute()
        self.graph = g
        return self

    def wait_idle(self):
        """Block until the last asynchronous replay has finished (h_in may be rewritten, h_out read)."""
        if self._done is not None:
            self._done.synchronize()
            self._done = None

    def __call__(self, sync=True):
        """h_in must already hold the inputs.  Returns the pinned result (valid until next call).
        sync=False only enqueues the replay: the result is valid after wait_idle() (or any later
        synchronisation of the stream), so the next call's host work overlaps this one's kernels."""
        if self.pre is not None:
            self._stage()
        self.graph.replay()
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
            self._done = None
        else:
            self._done = torch.cuda.Event()
            self._done.record(torch.cuda.current_stream(self.device))
        return self.h_out


class PendingScalar:
    """A host scalar produced by an asynchronous StagedGraph replay; float() waits for it."""

    def __init__(self, staged):
        self._staged = staged

    def __float__(self):
        self._staged.wait_idle()
        return float(self._staged.h_out[0])
