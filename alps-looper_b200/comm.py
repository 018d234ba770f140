"""Communicator adapters for the multi-GPU (imaginary-time slab) engine.

The C ABI (include/lq.h, lq_comm) asks the host for two collectives on DEVICE buffers:
an all-gather of the boundary cluster ids and an integer sum all-reduce of the open-cluster
partial sums.  `attach_torch_distributed` provides them with torch.distributed (NCCL over
NVLink/NVSwitch) on the engine's own CUDA stream; `LoopbackGroup` provides them for several
engines living in ONE process on ONE GPU (one Python thread per rank) so that the slab logic
can be tested without a multi-GPU box.
"""
import threading

import torch


class _DevBuf:
    """exposes a raw device pointer through __cuda_array_interface__ (zero copy)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 2}


_tensor_cache = {}


def device_tensor(ptr, nbytes, device):
    """zero-copy uint8 tensor over a raw device pointer (cached: the engine reuses its buffers)"""
    key = (int(ptr), int(nbytes), str(device))
    t = _tensor_cache.get(key)
    if t is None:
        if len(_tensor_cache) > 4096:
            _tensor_cache.clear()
        t = torch.as_tensor(_DevBuf(ptr, nbytes), device=device)
        _tensor_cache[key] = t
    return t


def attach_torch_distributed(engine, device, group=None):
    import torch.distributed as dist
    dev = torch.device("cuda", device)
    stream = torch.cuda.ExternalStream(engine.stream(), device=dev)

    def all_gather(ctx, send, recv, nbytes, strm):
        try:
            with torch.cuda.stream(stream):
                s = device_tensor(send, nbytes, dev)
                r = device_tensor(recv, nbytes * dist.get_world_size(group), dev)
                dist.all_gather_into_tensor(r, s, group=group)
            return 0
        except Exception as e:  # noqa: BLE001  (must not propagate through the C ABI)
            print("lq comm all_gather failed:", e, flush=True)
            return 1

    def all_reduce_i64(ctx, buf, count, strm):
        try:
            with torch.cuda.stream(stream):
                t = device_tensor(buf, count * 8, dev).view(torch.int64)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as e:  # noqa: BLE001
            print("lq comm all_reduce failed:", e, flush=True)
            return 1

    def send_recv(ctx, send, sbytes, dst, recv, rbytes, src, strm):
        try:
            with torch.cuda.stream(stream):
                ops = []
                if sbytes:
                    ops.append(dist.P2POp(dist.isend, device_tensor(send, sbytes, dev), dst, group))
                if rbytes:
                    ops.append(dist.P2POp(dist.irecv, device_tensor(recv, rbytes, dev), src, group))
                if ops:
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
            return 0
        except Exception as e:  # noqa: BLE001
            print("lq comm send_recv failed:", e, flush=True)
            return 1

    engine.set_comm(all_gather, all_reduce_i64, send_recv)


def attach_nccl(engine, rank, nranks, group=None):
    """The engine's own NCCL data plane (lq_comm_init): torch.distributed only carries the 128-byte
    unique id from rank 0 to the others; every collective of a step is then issued by the engine
    itself (ncclAllGather / ncclAllReduce on its stream), with no Python in the step."""
    import torch.distributed as dist
    import looper_b200 as lq
    box = [lq.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    engine.comm_init(box[0], rank, nranks)


class LoopbackGroup:
    """P engines in one process / one GPU, one thread each; collectives = device copies."""

    def __init__(self, nranks, device=0):
        self.n = nranks
        self.dev = torch.device("cuda", device)
        self.barrier = threading.Barrier(nranks)
        self.slots = [None] * nranks
        self.p2p = {}
        self.lock = threading.Lock()

    def attach(self, engine, rank):
        stream = torch.cuda.ExternalStream(engine.stream(), device=self.dev)

        def all_gather(ctx, send, recv, nbytes, strm):
            try:
                stream.synchronize()
                self.slots[rank] = device_tensor(send, nbytes, self.dev)
                self.barrier.wait()
                with torch.cuda.stream(stream):
                    r = device_tensor(recv, nbytes * self.n, self.dev)
                    for k in range(self.n):
                        r[k * nbytes:(k + 1) * nbytes].copy_(self.slots[k])
                stream.synchronize()
                self.barrier.wait()
                return 0
            except Exception as e:  # noqa: BLE001
                print("loopback all_gather failed:", e, flush=True)
                return 1

        def all_reduce_i64(ctx, buf, count, strm):
            try:
                stream.synchronize()
                self.slots[rank] = device_tensor(buf, count * 8, self.dev).view(torch.int64)
                self.barrier.wait()
                with torch.cuda.stream(stream):
                    total = torch.zeros(count, dtype=torch.int64, device=self.dev)
                    for k in range(self.n):
                        total += self.slots[k]
                stream.synchronize()
                self.barrier.wait()          # everybody has read all inputs
                with torch.cuda.stream(stream):
                    self.slots[rank].copy_(total)
                stream.synchronize()
                self.barrier.wait()
                return 0
            except Exception as e:  # noqa: BLE001
                print("loopback all_reduce failed:", e, flush=True)
                return 1

        def send_recv(ctx, send, sbytes, dst, recv, rbytes, src, strm):
            # (spatial cut: halo pages and ghost spins; every rank makes the same sequence of calls)
            try:
                stream.synchronize()
                self.p2p[(rank, dst)] = (send, sbytes)
                self.barrier.wait()
                ptr, nbytes = self.p2p[(src, rank)]
                assert nbytes == rbytes, (rank, src, nbytes, rbytes)
                if rbytes:
                    with torch.cuda.stream(stream):
                        device_tensor(recv, rbytes, self.dev).copy_(device_tensor(ptr, rbytes, self.dev))
                stream.synchronize()
                self.barrier.wait()
                return 0
            except Exception as e:  # noqa: BLE001
                print("loopback send_recv failed:", repr(e), flush=True)
                return 1

        engine.set_comm(all_gather, all_reduce_i64, send_recv)

    def run(self, fn):
        """fn(rank) in one thread per rank; returns the list of results (re-raises errors)."""
        out, err = [None] * self.n, [None] * self.n

        def body(r):
            try:
                torch.cuda.set_device(self.dev)
                out[r] = fn(r)
            except BaseException as e:  # noqa: BLE001
                err[r] = e
                self.barrier.abort()

        th = [threading.Thread(target=body, args=(r,)) for r in range(self.n)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for e in err:
            if e is not None:
                raise e
        return out
