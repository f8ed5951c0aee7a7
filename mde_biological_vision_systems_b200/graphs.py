"""CUDA-graph replay of the inference step (loader gather -> model -> losses) for fixed batch shapes.

The step is ~470 kernel launches of 2-300 us each; issued eagerly the GPU idles ~5 % of the time waiting for the host.
``GraphedStep`` captures one step into a CUDA graph (every kernel of this package is launched on the current stream
without host synchronisation, so it is capturable -- include/mde_b200.h) and replays it per batch: inputs are copied into
the static capture buffers, outputs are read from static tensors.  Shapes, modes and parameters must not change after
capture (in-place parameter updates are picked up only by the torch layers, the cached TF32 / folded weights are not).
"""
import torch


class GraphedStep:
    def __init__(self, fn, example_inputs, warmup=3):
        """fn(**inputs) -> tensor or tuple of tensors, all CUDA; example_inputs: dict of CUDA tensors (static buffers are
        cloned from them)."""
        self.static_in = {k: v.clone() for k, v in example_inputs.items()}
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):  # lazy initialisation (cuDNN plans, cached operands, smem attributes) outside capture
                for k, v in example_inputs.items():
                    self.static_in[k].copy_(v)
                self.fn(**self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        for k, v in example_inputs.items():
            self.static_in[k].copy_(v)
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = self.fn(**self.static_in)

    def __call__(self, **inputs):
        for k, v in inputs.items():
            self.static_in[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.static_out
