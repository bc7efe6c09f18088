"""Exchange steps of the D-sharded MoDE-conv path: halo planes, BatchNorm partial sums, gradient
sums between the ranks that hold the slabs of one volume.  Two interchangeable back ends behind one interface:

  * PeerComm  -- the product path on a B200 box: every rank maps the other ranks' exchange buffers (CUDA IPC over
                 NVLink / NVSwitch) and the exchange is a plain store into the neighbour's memory plus a counter
                 (mode_peer_put / mode_peer_wait / mode_peer_sum_slots, csrc/peer.cu).  No NCCL launch on the data
                 path, CUDA-graph capturable, deterministic summation order.
  * TorchComm -- the same steps through torch.distributed (NCCL on GPUs: the baseline PeerComm is measured against;
                 gloo in the CPU tests, and gloo with host staging when two test ranks share ONE GPU).

The reference has nothing to mirror here (fnet/fnet_model.py:40-44 is torch.nn.DataParallel).  Buffers that take part in
an exchange are allocated THROUGH the comm object (`alloc`) so that PeerComm can place them in the shared arena.
"""
import ctypes

import torch
import torch.distributed as dist

from . import lib as _lib


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class TorchComm:
    """Exchange steps as torch.distributed collectives.  `stage_host` copies CUDA tensors through the host around every
    collective (gloo cannot move CUDA memory point to point): only for the 2-ranks-on-1-GPU parity test."""

    kind = "torch"

    def __init__(self, group=None, stage_host=False):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.stage_host = stage_host
        self.log = None                  # set to a list to record (tag, numel) of every collective
        self.n_collectives = 0
        self._bufs = {}

    def alloc(self, name, shape, dtype, device):
        """Named persistent buffer, zero-filled once (halo planes at the global faces are never written)."""
        key = (name, tuple(shape), dtype, str(device))
        if key not in self._bufs:
            self._bufs[key] = torch.zeros(shape, dtype=dtype, device=device)
        return self._bufs[key]

    def _note(self, tag, t):
        self.n_collectives += 1
        if self.log is not None:
            self.log.append((tag, t.numel()))

    def all_reduce(self, t, tag=""):
        """In-place sum of `t` over the ranks."""
        if self.world == 1:
            return t
        self._note(tag, t)
        if self.stage_host and t.is_cuda:
            h = t.cpu()
            dist.all_reduce(h, group=self.group)
            t.copy_(h)
        else:
            dist.all_reduce(t, group=self.group)
        return t

    def halo_fill(self, ext, h, tag=""):
        """ext: [1, D + 2h, H, W, C] with the interior planes [h, D + h) written; fills planes [0, h) from the lower
        neighbour's last h interior planes and [D + h, D + 2h) from the upper neighbour's first h (global faces keep
        their zeros).  One grouped send/recv."""
        if self.world == 1:
            return ext
        assert ext.shape[0] == 1 and ext.is_contiguous()
        self._note(tag, ext[:, :2 * h])
        d = ext.shape[1] - 2 * h
        lo, hi = ext[0, :h], ext[0, d + h:]
        send_lo, send_hi = ext[0, h:2 * h], ext[0, d:d + h]
        stage = self.stage_host and ext.is_cuda
        if stage:
            lo_b, hi_b = torch.empty(lo.shape, dtype=lo.dtype), torch.empty(hi.shape, dtype=hi.dtype)
            send_lo, send_hi = send_lo.cpu(), send_hi.cpu()
        else:
            lo_b, hi_b = lo, hi
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, send_lo, self._peer(self.rank - 1), self.group))
            ops.append(dist.P2POp(dist.irecv, lo_b, self._peer(self.rank - 1), self.group))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, send_hi, self._peer(self.rank + 1), self.group))
            ops.append(dist.P2POp(dist.irecv, hi_b, self._peer(self.rank + 1), self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if stage:
            if self.rank > 0:
                lo.copy_(lo_b)
            if self.rank < self.world - 1:
                hi.copy_(hi_b)
        return ext

    def _peer(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    fused = False            # no kernel-fused exchange steps: callers use halo_fill / all_reduce

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)


class _RawCuda:
    """A raw device allocation presented through __cuda_array_interface__ so that torch can view it (no ownership)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class PeerComm:
    """Exchange steps as stores into peer memory (csrc/peer.cu).  Every rank allocates ONE arena, shares it through CUDA
    IPC and opens the arenas of the others; `alloc` carves named, identically laid out buffers out of it, so the address
    of a buffer on rank q is `peer_base[q] + offset`.  Counters live at the front of the arena."""

    kind = "peer"
    N_SIGNALS = 64

    def __init__(self, device, arena_bytes, group=None):
        assert dist.is_initialized(), "PeerComm needs an initialised process group to exchange the IPC handles"
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.n_collectives = 0
        self.log = None
        arena_bytes = (int(arena_bytes) + 4096 + 1023) // 1024 * 1024
        lib = _lib.load()
        self._base = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            # a dedicated, zero-filled cudaMalloc block: halo planes at the global faces are never written and must read as
            # the conv's zero padding
            _lib.check(lib.mode_peer_arena_alloc(arena_bytes, ctypes.byref(self._base), handle), "mode_peer_arena_alloc")
        self.arena = torch.as_tensor(_RawCuda(self._base.value, arena_bytes), device=self.device)
        handles = [None] * self.world
        dist.all_gather_object(handles, (bytes(handle.raw), self.device.index), group=group)
        self.peer_base = []                                       # address of rank q's arena in THIS process
        self._opened = []
        with torch.cuda.device(self.device):
            for q, (hb, dev_q) in enumerate(handles):
                if q == self.rank:
                    self.peer_base.append(self._base.value)
                    continue
                if dev_q != self.device.index:
                    _lib.check(lib.mode_peer_enable_access(int(dev_q)), "mode_peer_enable_access")
                ptr = ctypes.c_void_p()
                _lib.check(lib.mode_peer_arena_open(ctypes.create_string_buffer(hb, 64), ctypes.byref(ptr)),
                           "mode_peer_arena_open")
                self.peer_base.append(ptr.value)
                self._opened.append(ptr.value)
        self._off = 4096                                          # [0, 4096): counters
        self._bufs = {}
        self._sig_next = 0
        with torch.cuda.device(self.device):
            # local counters: expected values and launch tickets (never touched by a peer)
            self.local = torch.zeros(2 * self.N_SIGNALS, dtype=torch.int32, device=self.device)
        dist.barrier(group=group)

    # ---- arena bookkeeping (every rank performs the same sequence of allocations) ----
    def alloc(self, name, shape, dtype, device=None):
        if name in self._bufs:
            t = self._bufs[name][0]
            assert tuple(t.shape) == tuple(shape) and t.dtype == dtype, f"buffer {name} re-allocated with another shape"
            return t
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        off = (self._off + 1023) // 1024 * 1024
        if off + nbytes > self.arena.numel():
            raise RuntimeError(f"PeerComm arena exhausted allocating {name} ({nbytes} bytes)")
        self._off = off + nbytes
        t = self.arena[off:off + nbytes].view(dtype).view(shape)
        self._bufs[name] = (t, off)
        return t

    def _peer_ptr(self, q, t, byte_off=0):
        off = t.data_ptr() - self.arena.data_ptr()
        assert 0 <= off < self.arena.numel(), "tensor is not part of the exchange arena (allocate it with comm.alloc)"
        return self.peer_base[q] + off + byte_off

    def _signal(self, name):
        """Index of the (signal, expect, ticket) triple of a named exchange point."""
        key = ("sig", name)
        if key not in self._bufs:
            assert self._sig_next < self.N_SIGNALS
            self._bufs[key] = self._sig_next
            self._sig_next += 1
        return self._bufs[key]

    def _sig_ptrs(self, idx):
        local_sig = self.arena.data_ptr() + 4 * idx
        expect = self.local.data_ptr() + 4 * idx
        ticket = self.local.data_ptr() + 4 * (self.N_SIGNALS + idx)
        return local_sig, expect, ticket

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- exchange steps ----
    def halo_fill(self, ext, h, tag=""):
        """Push this rank's boundary planes into the neighbours' halo planes of the same buffer, then wait for theirs."""
        if self.world == 1:
            return ext
        lib = _lib.load()
        assert ext.shape[0] == 1 and ext.is_contiguous()
        self.n_collectives += 1
        d = ext.shape[1] - 2 * h
        plane = ext[0, 0].numel() * ext.element_size()
        idx = self._signal(("halo", tag))
        _, expect, ticket = self._sig_ptrs(idx)
        srcs, dsts, sigs = [], [], []
        if self.rank > 0:                                         # my first h interior planes -> lower neighbour's top halo
            srcs.append(ext.data_ptr() + h * plane)
            dsts.append(self._peer_ptr(self.rank - 1, ext, (d + h) * plane))
            sigs.append(self.peer_base[self.rank - 1] + 4 * idx)
        if self.rank < self.world - 1:                            # my last h interior planes -> upper neighbour's bottom halo
            srcs.append(ext.data_ptr() + d * plane)
            dsts.append(self._peer_ptr(self.rank + 1, ext, 0))
            sigs.append(self.peer_base[self.rank + 1] + 4 * idx)
        k = len(srcs)
        vp = ctypes.c_void_p * k
        _lib.check(lib.mode_peer_put(vp(*srcs), vp(*dsts), vp(*sigs), k, h * plane, ctypes.c_void_p(ticket), self._stream()),
                   "mode_peer_put")
        _lib.check(lib.mode_peer_wait(ctypes.c_void_p(self.arena.data_ptr() + 4 * idx), ctypes.c_void_p(expect), k,
                                      self._stream()), "mode_peer_wait")
        return ext

    def all_reduce(self, t, tag=""):
        """In-place sum over the ranks of a small vector (BatchNorm sums: fp64; gradients: fp32): every rank stores its
        vector into slot[rank] of every rank, then sums the slots in rank order (deterministic)."""
        if self.world == 1:
            return t
        lib = _lib.load()
        assert t.is_contiguous() and t.dtype in (torch.float32, torch.float64)
        self.n_collectives += 1
        n = t.numel()
        es = t.element_size()
        nbytes = (n * es + 15) // 16 * 16
        name = ("ar", tag, n, t.dtype)
        slots = self.alloc(name, (self.world, nbytes // es), t.dtype)
        idx = self._signal(name)
        _, expect, ticket = self._sig_ptrs(idx)
        direct = nbytes == n * es and t.data_ptr() % 16 == 0          # push from / sum into `t` itself: two launches in all
        if direct:
            src = t
        else:
            src = self.alloc(name + ("src",), (nbytes // es,), t.dtype)
            src[:n].copy_(t.reshape(-1))
        srcs = [src.data_ptr()] * self.world
        dsts = [self._peer_ptr(q, slots, self.rank * nbytes) for q in range(self.world)]
        sigs = [self.peer_base[q] + 4 * idx for q in range(self.world)]
        for q0 in range(0, self.world, 8):
            k = min(8, self.world - q0)
            vk = ctypes.c_void_p * k
            _lib.check(lib.mode_peer_put(vk(*srcs[q0:q0 + k]), vk(*dsts[q0:q0 + k]), vk(*sigs[q0:q0 + k]), k, nbytes,
                                         ctypes.c_void_p(ticket), self._stream()), "mode_peer_put")
        _lib.check(lib.mode_peer_sum_slots(_p(slots), self.world, nbytes // es, 1 if t.dtype == torch.float64 else 0,
                                           _p(src), ctypes.c_void_p(self.arena.data_ptr() + 4 * idx), ctypes.c_void_p(expect),
                                           ctypes.c_void_p(ticket), self._stream()), "mode_peer_sum_slots")
        if not direct:
            t.reshape(-1).copy_(src[:n])
        return t

    # ---- descriptors for exchange steps FUSED into kernels (mode_halo_push_t / mode_peer_push_t / mode_peer_gather_t) ----
    fused = True

    def halo_push_desc(self, ext, h, tag=""):
        """Descriptor for a kernel that writes the interior planes [h, D + h) of `ext` ([1, D + 2h, H, W, C], allocated with
        comm.alloc): its first / last h planes also go into the neighbours' halo planes.  Follow the kernel with
        halo_wait(desc).  Returns (ModeHaloPush, wait_args)."""
        assert ext.shape[0] == 1 and ext.is_contiguous()
        d = ext.shape[1] - 2 * h
        plane = ext[0, 0].numel() * ext.element_size()
        idx = self._signal(("halo", tag))
        _, expect, ticket = self._sig_ptrs(idx)
        hp = _lib.ModeHaloPush()
        k = 0
        if self.rank > 0:                                         # my first h interior planes -> lower neighbour's top halo
            hp.lo_dst = self._peer_ptr(self.rank - 1, ext, (d + h) * plane)
            hp.lo_signal = self.peer_base[self.rank - 1] + 4 * idx
            k += 1
        if self.rank < self.world - 1:                            # my last h interior planes -> upper neighbour's bottom halo
            hp.hi_dst = self._peer_ptr(self.rank + 1, ext, 0)
            hp.hi_signal = self.peer_base[self.rank + 1] + 4 * idx
            k += 1
        hp.bytes = h * plane
        hp.ticket = ticket
        return hp, (self.arena.data_ptr() + 4 * idx, expect, k)

    def halo_wait(self, wait_args):
        sig, expect, k = wait_args
        self.n_collectives += 1
        if k > 0:
            _lib.check(_lib.load().mode_peer_wait(ctypes.c_void_p(sig), ctypes.c_void_p(expect), k, self._stream()),
                       "mode_peer_wait")

    def reduce_desc(self, count, tag=""):
        """(ModePeerPush, ModePeerGather) for `count` doubles: the producing kernel's last block stores the rank's vector
        into slot[rank] of every rank; the consuming kernel waits for all of them and sums the slots in rank order."""
        name = ("fused_ar", tag, count)
        slots = self.alloc(name, (self.world, count), torch.float64)
        idx = self._signal(name)
        _, expect, ticket = self._sig_ptrs(idx)
        pp = _lib.ModePeerPush()
        pp.n = self.world
        for q in range(self.world):
            pp.dst[q] = self._peer_ptr(q, slots, self.rank * count * 8)
            pp.signal[q] = self.peer_base[q] + 4 * idx
        pp.ticket = ticket
        pg = _lib.ModePeerGather()
        pg.slots = slots.data_ptr()
        pg.world = self.world
        pg.signal = self.arena.data_ptr() + 4 * idx
        pg.expect = expect
        self.n_collectives += 1
        return pp, pg

    def barrier(self):
        dist.barrier(group=self.group)

    def close(self):
        """Unmap the peers' arenas and free our own (after a barrier: nobody may still be storing into it)."""
        if self._base is None:
            return
        lib = _lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            for p in self._opened:
                lib.mode_peer_arena_close(ctypes.c_void_p(p))
            dist.barrier(group=self.group)
            self.arena = None
            self._bufs = {}
            lib.mode_peer_arena_free(self._base)
        self._base = None
