"""Small helpers the hot path imports (reference: scoreperformer/utils/functions.py:12-34, 74-88)."""
from enum import Enum
from inspect import isfunction


def exists(val):
    return val is not None


def default(val, d):
    if exists(val):
        return val
    return d() if isfunction(d) else d


class equals:
    def __init__(self, val):
        self.val = val

    def __call__(self, x, *args, **kwargs):
        return x == self.val


def or_reduce(masks):
    head, *body = masks
    for rest in body:
        head = head | rest
    return head


class ExplicitEnum(str, Enum):
    """Enum with an explicit error message for missing values."""

    @classmethod
    def _missing_(cls, value):
        raise ValueError(f"{value} is not a valid {cls.__name__}, please select one of {list(cls._value2member_map_.keys())}")

    @classmethod
    def has_value(cls, value):
        return value in cls._value2member_map_

    @classmethod
    def list(cls):
        return list(map(lambda c: c.value, cls))


# ----------------------------------------------------------------------------- side-stream fork / join
# Loss-side work that nothing downstream waits for (MMD terms, classifier heads) consists of dozens of tiny kernels.  Run on a
# second CUDA stream it overlaps the large decoder kernels, in eager mode and as a parallel branch of the captured step graph;
# autograd replays each node on the stream of its forward, so the backward overlaps as well.  SPB_SIDE_STREAM=0 disables it.
import contextlib as _contextlib
import os as _os

_SIDE_STREAMS = {}
_TOUCHED = set()          # side streams that have run work since the last join_side_streams()


def join_side_streams() -> None:
    """Make the current stream wait for every side stream used since the last call.  Autograd replays each node on the stream
    of its forward, and nodes that write gradients in place (fused.DIRECT_GRAD) hand nothing back for the engine to order: the
    consumer of the gradient buffers (all-reduce, clip + AdamW) must wait for those streams itself."""
    if not _TOUCHED:
        return
    import torch
    cur = torch.cuda.current_stream()
    for s in list(_TOUCHED):
        if s != cur:
            cur.wait_stream(s)
    _TOUCHED.clear()


class SideBranch:
    def __init__(self, device, slot: int = 0):
        import torch
        self.torch = torch
        self.enabled = device.type == "cuda" and _os.environ.get("SPB_SIDE_STREAM", "1") != "0"
        self.side = None
        if self.enabled:
            key = (device.index if device.index is not None else torch.cuda.current_device(), slot)
            if key not in _SIDE_STREAMS:
                _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
                # gradients of shared parameters are intentionally produced on different streams (autograd syncs them)
                warn_off = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
                if warn_off is not None:
                    warn_off(False)
            self.side = _SIDE_STREAMS[key]
        self.used = False

    @_contextlib.contextmanager
    def run(self, *inputs):
        """Execute the body on the side stream, after everything issued so far on the current stream."""
        if not self.enabled:
            yield
            return
        cur = self.torch.cuda.current_stream()
        self.side.wait_stream(cur)
        for t in inputs:
            if t is not None and t.is_cuda:
                t.record_stream(self.side)
        self.used = True
        _TOUCHED.add(self.side)
        with self.torch.cuda.stream(self.side):
            yield

    def join(self, *outputs):
        """Make the current stream wait for the side work; `outputs` are side-stream tensors about to be used here."""
        if self.enabled and self.used:
            cur = self.torch.cuda.current_stream()
            cur.wait_stream(self.side)
            for t in outputs:
                if t is not None and t.is_cuda:
                    t.record_stream(cur)
            self.used = False
