"""Development aid (no GPU in the build container): runs bench.py's host-side control flow -- the timed loop, the
double-buffered end-to-end loop with its events and side streams, the precision / skipping modes, the JSON line, and the
CUDA-graph probe -- against a FAKE CUDA layer (no-op streams / events, a "graph" that re-runs the captured callable) and a
small CPU stand-in for the PVConv block.  It checks Python-level correctness of the judged artefact (no typo can reach the
round-end run unnoticed); it measures nothing and touches no kernel.  Not a product path.

    python tests/tools/dry_run_bench_on_cpu.py             # N=1: default line, graph probe, two --config lines
    python tests/tools/dry_run_bench_on_cpu.py --world 2   # N=2 over gloo: default line (+ strong scaling), --config lines
"""
import contextlib
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


class _Stream:
    def __init__(self, *a, **k):
        pass

    def wait_stream(self, s):
        pass

    def wait_event(self, e):
        pass

    def synchronize(self):
        pass


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class _Graph:
    fn = None

    def replay(self):
        self.fn()


@contextlib.contextmanager
def _ctx(*a, **k):
    yield


def _strip_device(fn):
    return lambda *a, **k: fn(*a, **{kk: v for kk, v in k.items() if kk != "device"})


def install_fake_cuda():
    cur = _Stream()
    torch.cuda.Stream, torch.cuda.Event, torch.cuda.CUDAGraph = _Stream, _Event, _Graph
    torch.cuda.current_stream = lambda *a, **k: cur
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.stream, torch.cuda.graph = _ctx, _ctx
    torch.Tensor.pin_memory = lambda self: self
    to = torch.Tensor.to

    def tensor_to(self, *a, **k):
        a = [x for x in a if not (isinstance(x, torch.device) and x.type == "cuda")]
        k = {kk: v for kk, v in k.items() if kk != "device"}
        return to(self, *a, **k) if (a or k) else self
    torch.Tensor.to = tensor_to
    torch.nn.Module.to = lambda self, *a, **k: self
    torch.empty, torch.zeros, torch.tensor = _strip_device(torch.empty), _strip_device(torch.zeros), _strip_device(torch.tensor)


class _Block(torch.nn.Module):
    """stand-in with the PVConv call contract: ((features, coords)) -> (features', coords)"""

    def __init__(self, cin, cout, k, r):
        super().__init__()
        self.conv, self.bn = torch.nn.Conv1d(cin, cout, 1), torch.nn.BatchNorm1d(cout)

    def forward(self, inputs):
        f, c = inputs
        return self.bn(self.conv(f)), c


def launch_world(world):
    """N>1: the same dry run as `world` processes over gloo (what the driver's scaling run does with torchrun + NCCL)"""
    import socket
    import subprocess
    rc = 0
    for what in ("ours", "s3dis_pvcnn", "shapenet_c0p25_train"):       # one rendezvous (port) per bench invocation
        sock = socket.socket()
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
        sock.close()
        procs = []
        for r in range(world):
            env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                       MASTER_PORT=str(port), DRY_RUN_WHAT=what)
            procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__)], env=env))
        rc = max([rc] + [p.wait(timeout=300) for p in procs])
    return rc


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--world":
        sys.exit(launch_world(int(sys.argv[2])))
    install_fake_cuda()
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        import torch.distributed as dist
        real_init = dist.init_process_group
        dist.init_process_group = lambda backend=None, **k: real_init("gloo")     # CPU tensors: gloo stands in for NCCL
    import bench
    import modules
    import pvcnn_b200.graphs as graphs
    import pvcnn_b200.parallel as parallel
    bench.B, bench.N, bench.C, bench.R = 2, 64, 8, 4
    bench.make_inputs.__defaults__ = (2,)
    bench.single_gpu_extras = lambda *a, **k: {}          # needs the real kernels
    modules.PVConv = _Block
    parallel.pin_process_to_gpu_numa_node = lambda i: None
    init = graphs.GraphedTrainStep.__init__

    def init_and_bind(self, *a, **k):
        init(self, *a, **k)
        self.graph.fn = self._step                         # the fake graph "replays" by re-running the captured step
    graphs.GraphedTrainStep.__init__ = init_and_bind
    what = os.environ.get("DRY_RUN_WHAT", "all")
    args = types.SimpleNamespace(precision="fp32", steps=6, warmup=3, scaling="weak", gpus=world)
    if what in ("all", "ours"):
        bench.run_ours(args)
    if what == "all":
        bench.run_graph_probe(args)
    # `bench.py --config <network>` (the children behind the `configs` sub-result): one inference and one training network
    from pvcnn_b200 import zoo

    class _Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.head = torch.nn.Conv1d(9, 50, 1)

        def forward(self, x):
            return self.head(x)
    specs = {"s3dis_pvcnn": dict(batch=2, points=32, channels=9, kind="s3dis", mode="eval"),
             "shapenet_c0p25_train": dict(batch=2, points=32, channels=9, kind="s3dis", mode="train")}
    zoo.build = lambda name: (_Net(), specs[name])
    zoo.synthetic_input = lambda spec, g, device="cpu", batch=None: torch.randn(spec["batch"], 9, spec["points"], generator=g)
    ginit = graphs.GraphedInference.__init__

    def ginit_and_bind(self, model, example_input, warmup=3):
        ginit(self, model, example_input, warmup)

        def again():
            with torch.no_grad():
                self.static_out = model(self.static_in)
        self.graph.fn = again
    graphs.GraphedInference.__init__ = ginit_and_bind
    for name in specs:
        if what in ("all", name):
            bench.run_config(types.SimpleNamespace(precision="fp32", steps=4, warmup=3, config=name))


if __name__ == "__main__":
    main()
