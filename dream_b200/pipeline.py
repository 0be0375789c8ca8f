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
    for the keypoints of batch i, so neither the PCIe copy nor the launch overhead sits on the critical path."""
    if network.network_config["architecture"]["output_heads"] != ["belief_maps"]:
        with torch.no_grad():
            for x in DevicePrefetcher(host_batches, network.device):
                yield network.inference(x)
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


def train_stream(network, host_batches):
    """host_batches yields (images, targets); yields the loss tensor of `network.train([x], t)`."""
    for x, t in DevicePrefetcher(host_batches, network.device):
        yield network.train([x], t)
