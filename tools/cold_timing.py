"""Per-launch time of a short kernel on COLD data (bench.py `other_kernels`, tools/perf_kernels.py).

The timing rules ask for inputs larger than L2 (or a flush) between timed iterations, and a 10-60 us kernel cannot be timed
launch by launch from Python (the host needs about as long per call).  So: K copies of the batch's packed states (together
>= 256 MiB, twice the 126 MB L2) and, for the row kernels, enough output buffers to exceed 256 MiB as well; launch k reads
copy k and writes buffer k mod B; the K launches are captured in ONE CUDA graph and the graph is timed per replay with CUDA
events on the replay stream.  The figure is the serialised per-launch time inside a stream (kernel + the gap to the next
launch).  State-changing calls get their copies restored between replays, outside the timed region.
"""
import math


def per_launch_ms(torch, eng, saved, call, out_spec=None, restore=False, replays=4, min_bytes=256 << 20):
    """call(out) runs ONE launch on eng.state (out: a buffer of out_spec = (shape, dtype), or None).  Returns the median over
    `replays` replays of (graph time / K)."""
    dev = eng.device
    state_bytes = saved.numel() * saved.element_size()
    K = max(8, math.ceil(min_bytes / state_bytes))
    copies = [saved.clone() for _ in range(K)]
    outs = [None]
    if out_spec is not None:
        shape, dtype = out_spec
        out_bytes = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
        outs = [torch.empty(shape, dtype=dtype, device=dev) for _ in range(max(2, math.ceil(min_bytes / out_bytes)))]
    orig = eng.state

    def body():
        for k, c in enumerate(copies):
            eng.state = c
            call(outs[k % len(outs)])

    try:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body()  # warm-up outside the capture (lazy allocations, module loading)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            body()
        times = []
        for _ in range(replays + 1):
            if restore:
                for c in copies:
                    c.copy_(saved)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            times.append(e0.elapsed_time(e1) / K)
    finally:
        eng.state = orig
    times = sorted(times[1:])  # the first replay pays the graph's upload
    return times[len(times) // 2]
