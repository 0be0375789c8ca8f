"""Checkpoint / resume bookkeeping around the training hot path.  Host Python only.

On-disk contract = what `scripts/train_network.py` writes and reads, so a run started with the reference can be resumed
here and vice versa:
  * per epoch N: `epoch_{N}.pth` (`model.state_dict()`, keys with the `module.` prefix) + `epoch_{N}.yaml` (the whole
    network_config incl. `training.results`) -- train_network.py:656-659 through `DreamNetwork.save_network`;
  * `best_network.{pth,yaml}` whenever the mean validation loss improves (:612-620);
  * the rolling pickle `training_log_e{N}.pkl`, renamed to `training_log.pkl` when training ends (:641-665);
  * resume (:66-147, :326-407): newest `epoch_*.pth`, best validation loss from `best_network.yaml`, random seed from the
    training log, and a list of config keys that must not have changed.
One file is added that the reference does not have: `optim_epoch_{N}.pt` with the optimizer state (the reference
restarts Adam's moments on every resume).  It is optional on both sides: absent -> the reference's behaviour.  The name
is chosen so that the reference's resume scan (train_network.py:70-85: every file that starts with "epoch" and ends
with ".pth" is taken for a weights file) can never pick it up.

`AsyncCheckpointWriter` takes the device -> host copy and the `torch.save` off the training thread: the state dict is
snapshotted into pinned host buffers on a side stream (the next step's kernels are not blocked) and written by a worker
thread; the files are the same.
"""
import os
import pickle
import queue
import re
import threading
import time

import torch

from .network import dump_yaml_config, load_yaml_config

# keys scripts/train_network.py:338-391 asserts unchanged between the saved config and the new command line
_TOP_LEVEL_KEYS = ("data_path", "manipulator", "architecture")
_TRAINING_CONFIG_KEYS = ("training_data_fraction", "validation_data_fraction", "batch_size", "data_augmentation",
                         "worker_size", "optimizer", "image_preprocessing", "image_raw_resolution",
                         "net_input_resolution")
_EPOCH_RE = re.compile(r"^epoch_(\d+)\.pth$")


def list_epoch_checkpoints(output_dir):
    """[(epoch, weights_path, config_path)] sorted by epoch, newest first (train_network.py:70-85)."""
    found = []
    for name in os.listdir(output_dir):
        m = _EPOCH_RE.match(name)
        if m:
            path = os.path.join(output_dir, name)
            found.append((int(m.group(1)), path, path[:-4] + ".yaml"))
    found.sort(key=lambda t: t[0], reverse=True)
    return found


def optimizer_state_path(output_dir, epoch):
    return os.path.join(output_dir, "optim_epoch_{}.pt".format(epoch))


def assert_resume_consistent(saved_config, new_config):
    """The consistency checks of train_network.py:338-391; a missing key on BOTH sides is not a difference."""
    for key in _TOP_LEVEL_KEYS:
        assert saved_config.get(key) == new_config.get(key), \
            'Cannot resume: "{}" differs from the checkpoint\'s.'.format(key)
    saved_t = saved_config.get("training", {}).get("config", {})
    new_t = new_config.get("training", {}).get("config", {})
    for key in _TRAINING_CONFIG_KEYS:
        assert saved_t.get(key) == new_t.get(key), \
            'Cannot resume: training config "{}" differs from the checkpoint\'s.'.format(key)


def read_best_validation_loss(output_dir):
    path = os.path.join(output_dir, "best_network.yaml")
    assert os.path.exists(path), "Could not determine the best validation loss."          # train_network.py:98-100
    return float(load_yaml_config(path)["training"]["results"]["validation_loss"]["mean"])


def load_training_log(output_dir, start_epoch, now=None):
    """The log juggling of train_network.py:108-147: `training_log.pkl` (a finished run) becomes
    `training_log_e{start_epoch}.pkl` again; start time and the resumed epoch are appended."""
    final_path = os.path.join(output_dir, "training_log.pkl")
    epoch_path = os.path.join(output_dir, "training_log_e{}.pkl".format(start_epoch))
    if os.path.exists(final_path):
        with open(final_path, "rb") as f:
            log = pickle.load(f)
        os.rename(final_path, epoch_path)
    elif os.path.exists(epoch_path):
        with open(epoch_path, "rb") as f:
            log = pickle.load(f)
    else:
        assert False, "Could not determine training log file to resume."
    if not isinstance(log["start_time"], list):
        log["start_time"] = [log["start_time"]]
    log["start_time"].append(time.time() if now is None else now)
    log.setdefault("epochs_resumed", []).append(start_epoch + 1)
    return log


class ResumePoint:
    """What `find_resume_point` hands back: where to continue and with what."""

    def __init__(self, start_epoch, weights_path, config_path, network_config, best_valid_loss, train_log,
                 optimizer_path):
        self.start_epoch = start_epoch
        self.weights_path = weights_path
        self.config_path = config_path
        self.network_config = network_config          # the SAVED config: "use this one instead" (:393)
        self.best_valid_loss = best_valid_loss
        self.train_log = train_log
        self.random_seed = train_log["random_seed"] if train_log is not None else None
        self.optimizer_path = optimizer_path          # None when the checkpoint has no optimizer state


def find_resume_point(output_dir, new_network_config=None, total_epochs=None, load_log=True):
    checkpoints = list_epoch_checkpoints(output_dir)
    assert checkpoints, 'No "epoch_*.pth" checkpoint in "{}".'.format(output_dir)
    start_epoch, weights_path, config_path = checkpoints[0]
    if total_epochs is not None:
        assert start_epoch < total_epochs, "Network is already trained for the number of requested epochs."
    best = read_best_validation_loss(output_dir)
    saved = load_yaml_config(config_path)
    if new_network_config is not None:
        assert_resume_consistent(saved, new_network_config)
    log = load_training_log(output_dir, start_epoch) if load_log else None
    optim_path = optimizer_state_path(output_dir, start_epoch)
    return ResumePoint(start_epoch, weights_path, config_path, saved, best, log,
                       optim_path if os.path.exists(optim_path) else None)


def restore(network, resume_point, map_location=None):
    """Weights into `network.model` (train_network.py:404-406), then `enable_training()` and, when the checkpoint
    carries it, the optimizer state."""
    network.model.load_state_dict(torch.load(resume_point.weights_path, map_location=map_location))
    network.enable_training()
    if resume_point.optimizer_path is not None and network.optimizer is not None:
        network.optimizer.load_state_dict(torch.load(resume_point.optimizer_path, map_location=map_location))
    return network


def _write_log(output_dir, epoch, train_log, previous_epoch):
    path = os.path.join(output_dir, "training_log_e{}.pkl".format(epoch))
    with open(path, "wb") as f:
        pickle.dump(train_log, f)
    if previous_epoch is not None and previous_epoch != epoch:
        old = os.path.join(output_dir, "training_log_e{}.pkl".format(previous_epoch))
        if os.path.exists(old):
            os.remove(old)
    return path


def save_epoch(network, output_dir, epoch, train_log=None, previous_epoch=None, is_best=False, save_optimizer=True,
               writer=None):
    """One epoch's files (train_network.py:612-659).  `writer`: an `AsyncCheckpointWriter` to take it off-thread."""
    os.makedirs(output_dir, exist_ok=True)
    names = ["epoch_{}".format(epoch)] + (["best_network"] if is_best else [])
    if train_log is not None:
        _write_log(output_dir, epoch, train_log, previous_epoch)
    optim_state = network.optimizer.state_dict() if (save_optimizer and network.optimizer is not None) else None
    if writer is not None:
        writer.submit(network.model.state_dict(), network.network_config, output_dir, names, optim_state,
                      optim_path=optimizer_state_path(output_dir, epoch))
        return
    for name in names:
        network.save_network(output_dir, name, overwrite=True)
    if optim_state is not None:
        torch.save(optim_state, optimizer_state_path(output_dir, epoch))


def finish_training(output_dir, last_epoch):
    """`training_log_e{last}.pkl` -> `training_log.pkl` (train_network.py:661-665)."""
    src = os.path.join(output_dir, "training_log_e{}.pkl".format(last_epoch))
    if os.path.exists(src):
        os.rename(src, os.path.join(output_dir, "training_log.pkl"))


def _snapshot(obj, stream):
    """Deep copy of a (nested) state dict with every tensor in host memory.  A CUDA tensor is first CLONED on the
    current (training) stream -- `state_dict()` / `optimizer.state_dict()` alias the live parameters, Adam moments and
    BatchNorm buffers, which the next step updates in place -- and the clone is then copied to a pinned buffer on
    `stream`; the caller makes `stream` wait for the clones before the copies start."""
    if torch.is_tensor(obj):
        if obj.is_cuda:
            frozen = obj.detach().clone()                         # ordered before the next step's in-place updates
            host = torch.empty(obj.shape, dtype=obj.dtype, device="cpu", pin_memory=True)
            return _Pending(frozen, host)
        return obj.detach().clone()
    if isinstance(obj, dict):
        return type(obj)((k, _snapshot(v, stream)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        return type(obj)(_snapshot(v, stream) for v in obj)
    return obj


class _Pending:
    """A device clone waiting for its device -> host copy."""
    __slots__ = ("dev", "host")

    def __init__(self, dev, host):
        self.dev, self.host = dev, host


def _drain(obj, stream):
    """Second pass of a snapshot: issue the device -> host copies on `stream` and return the structure with the
    pinned host tensors in place of the `_Pending` markers."""
    if isinstance(obj, _Pending):
        with torch.cuda.stream(stream):
            obj.host.copy_(obj.dev, non_blocking=True)
            obj.dev.record_stream(stream)                         # the clone may be freed only after the copy
        return obj.host
    if isinstance(obj, dict):
        return type(obj)((k, _drain(v, stream)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        return type(obj)(_drain(v, stream) for v in obj)
    return obj


class AsyncCheckpointWriter:
    """Writes checkpoints on a worker thread.  `submit` freezes the tensors with device-side clones on the training
    stream (the state after the step that was just queued; later in-place updates cannot tear it), copies the clones to
    pinned host buffers on a side stream and returns; the worker waits for the copies and writes the same files
    `DreamNetwork.save_network` writes."""

    def __init__(self):
        self._q = queue.Queue()
        self._err = None
        self._stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def submit(self, state_dict, network_config, output_dir, names, optim_state=None, optim_path=None):
        self._raise_pending()
        done = None
        # 1. freeze: device-side clones on the training stream (cheap, ordered before any later in-place update)
        weights = _snapshot(state_dict, self._stream)
        optim = _snapshot(optim_state, self._stream) if optim_state is not None else None
        if self._stream is not None:
            # 2. copy out on the side stream once the clones exist; the training stream never waits for this
            self._stream.wait_stream(torch.cuda.current_stream())
            weights = _drain(weights, self._stream)
            optim = _drain(optim, self._stream) if optim is not None else None
            done = torch.cuda.Event()
            done.record(self._stream)
        config = pickle.loads(pickle.dumps(network_config))      # the caller keeps mutating training.results
        if optim is not None and optim_path is None:
            m = re.match(r"^epoch_(\d+)$", names[0])
            optim_path = optimizer_state_path(output_dir, int(m.group(1))) if m else \
                os.path.join(output_dir, "optim_" + names[0] + ".pt")
        self._q.put((weights, optim, config, output_dir, list(names), done, optim_path))

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                self._q.task_done()
                return
            try:
                weights, optim, config, output_dir, names, done, optim_path = item
                if done is not None:
                    done.synchronize()
                for name in names:
                    dump_yaml_config(config, os.path.join(output_dir, name + ".yaml"))
                    torch.save(weights, os.path.join(output_dir, name + ".pth"))
                if optim is not None:
                    torch.save(optim, optim_path)
            except Exception as e:  # surfaced on the next submit / wait
                self._err = e
            finally:
                self._q.task_done()

    def _raise_pending(self):
        if self._err is not None:
            err, self._err = self._err, None
            raise err

    def wait(self):
        self._q.join()
        self._raise_pending()

    def close(self):
        self.wait()
        self._q.put(None)
        self._thread.join(timeout=10)
