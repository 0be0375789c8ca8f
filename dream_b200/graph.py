"""Whole-step CUDA graphs of the inference path (SURVEY.md 8b `dreamb200_network_fwd`: the "whole-graph variant").

`DreamNetwork.inference` is ~35 kernel launches (23 conv layers as 30 launches, pools, peak extraction, the decision
table).  Eagerly, every launch pays a ctypes crossing, output allocation and -- for the tensor-core kernels -- the
host-side encoding of three or four `CUtensorMap`s; at B = 128 that hides behind 14 ms of GPU work, at B = 1 (the
`keypoints_from_image` / ROS path, dream/network.py:423-499, scripts/launch_dream_ros.py:223-256) it IS the latency.
A captured graph replays the same kernels with their tensor maps and arguments frozen: one driver call per step, no
Python between launches, no inter-launch host gaps.

The graph is tied to (input buffer, shape, dtype, weights version): `InferenceGraph` owns a static input buffer (or
adopts the caller's) and static outputs; `DreamNetwork.capture_inference` / `inference_graphed` manage a small cache.
Capture uses torch's graph-aware allocator pool, so the intermediate activations of a graph are private to it.
"""
import torch


class InferenceGraph:
    """`DreamNetwork.inference_device` for one fixed input buffer, captured once and replayed.

    g = InferenceGraph(network, x)         # x: CUDA input batch; adopt=True replays straight from THIS buffer
    belief, kps = g(x_new)                 # copies x_new into the static input (skipped when it IS the buffer), replays
    The returned tensors are the graph's static outputs: valid until the next replay (clone to keep)."""

    def __init__(self, network, example, adopt=False, warmup=2):
        assert example.is_cuda, "InferenceGraph needs a CUDA input"
        self.network = network
        self.static_in = example if adopt else example.clone()
        self.key = graph_key(network, example)
        cur = torch.cuda.current_stream(example.device)
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                       # packs the weights, sets kernel attributes, warms the allocator
                network._inference_device_eager(self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize(example.device)
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.belief, self.kps = network._inference_device_eager(self.static_in)
        self.kernels_per_replay = _lib.launch_count() - n0      # libdreamb200 kernels inside one replay
        self.replays = 0

    def __call__(self, x=None):
        if x is not None and x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.belief, self.kps


def graph_key(network, x):
    """Everything a captured step depends on besides the input VALUES."""
    model = network.model.module
    ver = tuple((t.data_ptr(), t._version) for t in list(model.parameters()) + list(model.buffers()))
    return (tuple(x.shape), x.dtype, model.training, network.use_belief_peak_scores,
            float(network.belief_peak_next_best_score), hash(ver))
