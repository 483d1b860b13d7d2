"""dist — the data-parallel exchange: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch; gloo in the
CPU unit tests) used as plumbing for the single flat all-reduce per step."""
from __future__ import annotations

import os

import torch
import torch.distributed as td


def is_initialized() -> bool:
    return td.is_available() and td.is_initialized()


def world_size() -> int:
    return td.get_world_size() if is_initialized() else 1


def rank() -> int:
    return td.get_rank() if is_initialized() else 0


def init_from_env(backend: str = "nccl"):
    """Initialise from torchrun's RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (no-op for a single process)."""
    if is_initialized() or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    td.init_process_group(backend=backend)


def all_reduce_sum(buf: torch.Tensor):
    if world_size() > 1:
        td.all_reduce(buf, op=td.ReduceOp.SUM)


def broadcast(buf: torch.Tensor, src: int = 0):
    if world_size() > 1:
        td.broadcast(buf, src=src)


# ---- a communicator of our own, for collectives that live INSIDE the step's CUDA graph ---------------------------------------
# torch.distributed's process group wraps every collective in its own stream / event / watchdog bookkeeping; capturing that into
# the training-step graph hung at the first replay (round 2, first session).  ncclAllReduce itself is graph-capturable: it only
# enqueues a kernel on the stream it is given.  GraphComm opens a second communicator on the NCCL library torch has already
# loaded (same ranks, id broadcast through the process group) and enqueues on torch's CURRENT stream — under
# torch.cuda.graph() that is the capture stream.
import ctypes as _C

_NCCL_FLOAT32, _NCCL_SUM = 7, 0


class _NcclUniqueId(_C.Structure):
    _fields_ = [("internal", _C.c_byte * 128)]


def _loaded_nccl_path():
    with open("/proc/self/maps") as f:
        for line in f:
            if "libnccl.so" in line:
                return line.split()[-1]
    return None


class GraphComm:
    def __init__(self):
        if not (is_initialized() and td.get_backend() == "nccl"):
            raise RuntimeError("GraphComm needs torch.distributed initialised with the nccl backend")
        path = _loaded_nccl_path()
        if path is None:
            raise RuntimeError("libnccl is not loaded in this process")
        self.lib = _C.CDLL(path)
        self.lib.ncclGetErrorString.restype = _C.c_char_p
        self.lib.ncclCommInitRank.argtypes = [_C.POINTER(_C.c_void_p), _C.c_int, _NcclUniqueId, _C.c_int]
        self.lib.ncclAllReduce.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_size_t, _C.c_int, _C.c_int, _C.c_void_p, _C.c_void_p]
        uid = _NcclUniqueId()
        if rank() == 0:
            self._check(self.lib.ncclGetUniqueId(_C.byref(uid)), "ncclGetUniqueId")
        t = torch.tensor(list(bytes(uid.internal)), dtype=torch.uint8, device="cuda")
        td.broadcast(t, src=0)
        raw = bytes(t.cpu().tolist())
        _C.memmove(_C.byref(uid), raw, 128)
        self.comm = _C.c_void_p()
        torch.cuda.synchronize()
        self._check(self.lib.ncclCommInitRank(_C.byref(self.comm), world_size(), uid, rank()), "ncclCommInitRank")
        warm = torch.zeros(256, dtype=torch.float32, device="cuda")   # connection set-up happens at the first collective: not under capture
        self.all_reduce_sum(warm)
        torch.cuda.synchronize()

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed: {self.lib.ncclGetErrorString(rc).decode()}")

    def all_reduce_sum(self, buf: torch.Tensor):
        """in place, fp32, on torch's current stream (the capture stream under torch.cuda.graph)"""
        if buf.dtype != torch.float32 or not buf.is_contiguous():
            raise ValueError("GraphComm.all_reduce_sum: contiguous float32 tensors only")
        st = torch.cuda.current_stream().cuda_stream
        self._check(self.lib.ncclAllReduce(buf.data_ptr(), buf.data_ptr(), buf.numel(), _NCCL_FLOAT32, _NCCL_SUM, self.comm,
                                           _C.c_void_p(st)), "ncclAllReduce")


_graph_comm = None


def graph_comm():
    """the process-wide GraphComm when VQB_DP_INGRAPH=1 asks for it and it can be had (NCCL backend), else None"""
    global _graph_comm
    if _graph_comm is None:
        # opt-in: measured on 2 and 8 B200 the captured collective is correct (tools/dp_check.py: ranks bit-identical) but no faster
        # than the all-reduce between two graph replays (2 GPUs: 7.60 vs 7.57 ms; 8 GPUs: 7.63 vs 7.57 ms per step)
        if os.environ.get("VQB_DP_INGRAPH", "0") != "1" or not is_initialized() or td.get_backend() != "nccl":
            _graph_comm = False
        else:
            try:
                _graph_comm = GraphComm()
            except Exception as e:  # noqa: BLE001 - any failure here means: keep the collective between two graph replays
                print(f"vqvae_b200.dist: in-graph collective unavailable ({e}); the all-reduce stays between two graph replays", flush=True)
                _graph_comm = False
    return _graph_comm or None
