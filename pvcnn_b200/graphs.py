"""CUDA-graph replay of an inference forward pass.

Whole networks built from these modules launch 120-500 kernels per forward (S3DIS PVCNN 129, PVCNN++ 498, Frustum-PVCNN 180),
each a few microseconds of GPU work at the reference's batch sizes, so eager execution is bound by the host's launch rate,
not by the GPU.  Every kernel of the native path is launched on torch's current stream through the C ABI and none of them
needs a host round trip (the activity lists, FPS, ball query and the `logits_mask` resampling all stay on the device), so
an eval-mode forward can be captured once into a CUDA graph and replayed with new inputs.

    g = GraphedInference(model, example_input)     # eval mode, fixed shapes
    out = g(new_input)                             # copy-in, one graph launch; `out` is overwritten by the next call

Limits: shapes, precision mode and env knobs are frozen at capture time; the `logits_mask` seed drawn during capture is
baked into the graph (the resampling is then the same every replay).

`GraphedTrainStep` does the same for one forward + backward of a block (fixed shapes): the fused PVConv step is 54 short
launches plus autograd bookkeeping, ~2 ms of host work for ~2 ms of GPU work, so a loop that also moves host buffers every
step becomes host bound.  Parameters are read at replay time (an optimizer step between replays is seen); gradients land in
the tensors that existed / were created at capture (`p.grad`, or the flat bucket of pvcnn_b200/parallel.py).
Status: written after the round's GPU budget was spent; exercised only by `bench.py`'s contained `modes.cuda_graph` probe
(a child process that also checks the replay against an eager step), not by a parity test.
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


class GraphedTrainStep:
    """forward + backward of `module((features, coords))` against a fixed output gradient, captured once.

        g = GraphedTrainStep(block, features, coords, grad_out, bucket=None)
        out, grad_features = g(features, coords, grad_out)      # copy-in, one graph launch; results are overwritten
                                                                # by the next call; parameter gradients are in p.grad
    """

    def __init__(self, module, features, coords, grad_out, bucket=None, warmup=3):
        if not module.training:
            raise RuntimeError("GraphedTrainStep captures a training step; call module.train() first")
        self.module, self.bucket = module, bucket
        self.f = features.detach().clone().requires_grad_(True)
        self.c = coords.detach().clone()
        self.go = grad_out.detach().clone()
        self.out = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):      # shared scratch, per-device flags and autograd's lazily built state exist before capture
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step()

    def _step(self):
        if self.bucket is not None:
            self.bucket.zero()
        else:
            for p in self.module.parameters():
                p.grad = None            # during capture the new gradient tensors come from the graph's pool and stay put
        self.f.grad = None
        out, _ = self.module((self.f, self.c))
        out.backward(self.go)
        if self.bucket is not None:
            self.bucket.finish()
        self.out = out.detach()

    def __call__(self, features=None, coords=None, grad_out=None):
        with torch.no_grad():
            if features is not None:
                self.f.copy_(features, non_blocking=True)
            if coords is not None:
                self.c.copy_(coords, non_blocking=True)
            if grad_out is not None:
                self.go.copy_(grad_out, non_blocking=True)
        self.graph.replay()
        return self.out, self.f.grad
