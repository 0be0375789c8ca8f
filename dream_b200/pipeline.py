"""Host -> device input pipeline for the hot path.

The reference's loops copy each batch synchronously right before using it
(`sample["image_rgb_input"].cuda()`, dream/analysis.py:209, scripts/train_network.py:494-497), so the
PCIe transfer (246 MB per 128-frame fp32 batch) is serialised with the network.  `DevicePrefetcher`
issues the copy of batch i+1 on a side stream while batch i computes; `inference_stream` /
`train_stream` wrap `DreamNetwork.inference` / `.train` around it.  Batches should live in pinned
host memory (`DataLoader(pin_memory=True)`) for the copies to be asynchronous.
"""
import torch


class DevicePrefetcher:
    """Iterates host batches (a tensor or a tuple/list of tensors) and yields them on `device`,
    with the next batch's H2D copy already in flight on a dedicated stream."""

    def __init__(self, host_batches, device):
        self.it = iter(host_batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.next = None
        self._preload()

    def _to_device(self, item):
        if torch.is_tensor(item):
            return item.to(self.device, non_blocking=True)
        return type(item)(self._to_device(t) for t in item)

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self.next = None
            return
        with torch.cuda.stream(self.stream):
            self.next = self._to_device(host)

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream)                 # batch i has landed
        batch = self.next
        for t in ([batch] if torch.is_tensor(batch) else batch):
            t.record_stream(cur)                     # allocated on the copy stream, consumed on the compute stream
        self._preload()                              # batch i+1 starts copying while batch i computes
        return batch


def inference_stream(network, host_batches):
    """Yields `[belief_maps (cuda), keypoints (cpu fp32 [B,K,2])]` -- what `network.inference(x)` returns -- for
    every host batch.  Software-pipelined by one batch: the kernels of batch i+1 are queued before the host waits
    for the keypoints of batch i, so neither the PCIe copy nor the launch overhead sits on the critical path.

    With `network.use_cuda_graphs` (default) every batch shape gets a PAIR of captured CUDA graphs (dream_b200.graph)
    with their own static input buffers: batch i+1 is copied host -> device straight into the idle graph's input
    while graph i replays, and a step is one driver call instead of ~35 (vgg) or ~340 (resnet) Python launches --
    the resnet paths were host-bound without it.  A batch of another shape (the ragged last one) gets its own pair
    on first sight."""
    if network.network_config["architecture"]["output_heads"] != ["belief_maps"]:
        with torch.no_grad():
            for x in DevicePrefetcher(host_batches, network.device):
                yield network.inference(x)
        return
    if getattr(network, "use_cuda_graphs", False):
        yield from _inference_stream_graphed(network, host_batches)
        return
    pending = None
    with torch.no_grad():
        for x in DevicePrefetcher(host_batches, network.device):
            belief, kps_dev = network.inference_device(x)
            kps_host = torch.empty(kps_dev.shape, dtype=torch.float32).pin_memory()
            kps_host.copy_(kps_dev, non_blocking=True)          # float64 -> float32 conversion happens on the device
            done = torch.cuda.Event()
            done.record()
            if pending is not None:
                pending[2].synchronize()
                yield [pending[0], pending[1]]
            pending = (belief, kps_host, done)
        if pending is not None:
            pending[2].synchronize()
            yield [pending[0], pending[1]]


def _inference_stream_graphed(network, host_batches):
    device = network.device
    cur = torch.cuda.current_stream(device)
    copy_stream = torch.cuda.Stream(device=device)
    from .graph import graph_key
    # graph pairs live on the network (captured once per input shape and weights version, reused by later streams):
    # key -> {"graphs": [g0, g1], "free": [event or None] * 2, "next": 0}
    pairs = network.__dict__.setdefault("_stream_graphs", {})
    pending = None
    with torch.no_grad():
        for host_x in host_batches:
            key = graph_key(network, host_x)
            slot = pairs.get(key)
            if slot is None:
                while len(pairs) >= 3:
                    pairs.pop(next(iter(pairs)))
                example = host_x.to(device)
                slot = pairs[key] = {"graphs": [network.capture_inference(example) for _ in range(2)],
                                     "free": [None, None], "next": 0}
            i = slot["next"]
            slot["next"] ^= 1
            g = slot["graphs"][i]
            if slot["free"][i] is not None:
                copy_stream.wait_event(slot["free"][i])          # the replay that last read this input buffer is done
            with torch.cuda.stream(copy_stream):
                g.static_in.copy_(host_x, non_blocking=True)
                landed = torch.cuda.Event()
                landed.record(copy_stream)
            cur.wait_event(landed)
            belief_static, kps_dev = g()                         # one driver call: the whole step
            slot["free"][i] = torch.cuda.Event()
            slot["free"][i].record(cur)
            belief = belief_static.clone()                       # the graph's output buffer is reused two batches on
            kps_host = torch.empty(kps_dev.shape, dtype=torch.float32).pin_memory()
            kps_host.copy_(kps_dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(cur)
            if pending is not None:
                pending[2].synchronize()
                yield [pending[0], pending[1]]
            pending = (belief, kps_host, done)
        if pending is not None:
            pending[2].synchronize()
            yield [pending[0], pending[1]]


def train_stream(network, host_batches):
    """host_batches yields (images, targets); yields the loss tensor of `network.train([x], t)`."""
    for x, t in DevicePrefetcher(host_batches, network.device):
        yield network.train([x], t)
