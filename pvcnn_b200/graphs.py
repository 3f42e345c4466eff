"""CUDA-graph replay of an inference forward pass.

Whole networks built from these modules launch 120-500 kernels per forward (S3DIS PVCNN 129, PVCNN++ 498, Frustum-PVCNN 180),
each a few microseconds of GPU work at the reference's batch sizes, so eager execution is bound by the host's launch rate,
not by the GPU.  Every kernel of the native path is launched on torch's current stream through the C ABI and none of them
needs a host round trip (the activity lists, FPS, ball query and the `logits_mask` resampling all stay on the device), so
an eval-mode forward can be captured once into a CUDA graph and replayed with new inputs.

    g = GraphedInference(model, example_input)     # eval mode, fixed shapes
    out = g(new_input)                             # copy-in, one graph launch; `out` is overwritten by the next call

Limits: shapes, precision mode and env knobs are frozen at capture time; the `logits_mask` seed drawn during capture is
baked into the graph (the resampling is then the same every replay); training (which needs autograd) is not graphed.
"""
import torch


def _clone(x):
    if isinstance(x, dict):
        return {k: _clone(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_clone(v) for v in x)
    return x.clone() if isinstance(x, torch.Tensor) else x


def _copy(dst, src):
    if isinstance(dst, dict):
        for k in dst:
            _copy(dst[k], src[k])
    elif isinstance(dst, (list, tuple)):
        for d, s in zip(dst, src):
            _copy(d, s)
    elif isinstance(dst, torch.Tensor):
        dst.copy_(src, non_blocking=True)


class GraphedInference:
    def __init__(self, model, example_input, warmup=3):
        if model.training:
            raise RuntimeError("GraphedInference captures an eval-mode forward; call model.eval() first")
        self.model = model
        self.static_in = _clone(example_input)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):      # lazily created buffers (shared scratch, per-device flags) exist before capture
                model(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = model(self.static_in)

    def __call__(self, x):
        _copy(self.static_in, x)
        self.graph.replay()
        return self.static_out
