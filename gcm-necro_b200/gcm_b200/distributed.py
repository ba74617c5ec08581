"""One process per GPU: the DataBus role (gcm/system/DataBus.cpp) on torch.distributed.

``DataBus.create_custom_types`` exchanges the halo index lists once (DataBus.cpp:172-498),
``sync_nodes_startup`` is the first ``DataBus::sync_nodes`` on host memory (coords + 13 values,
main.cpp:127), ``do_next_step`` is ``TetrMeshSet::do_next_step`` (TetrMeshSet.cpp:101-184) with the
cross-rank part of ``sync_nodes`` done as one grouped NCCL send/recv per neighbour before each
stage, stream-ordered on the set's CUDA stream (no host synchronisation inside a step).  By default
the set owns the communicator (``init_native_comm``: the id travels over torch.distributed once)
and a step is one library call; ``GCM_B200_TORCH_HALO=1`` keeps the per-stage exchange in Python
through ``torch.distributed`` P2P ops on caller-owned buffers (``gcm_b200_halo_pack/unpack``).
With the gloo backend (CPU tests) only the planning / start-up exchange is available.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class DataBus:
    def __init__(self, mesh_set, group=None):
        self.ms = mesh_set
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.peers = []          # ranks this process exchanges halo nodes with
        self.n_recv = {}
        self.n_send = {}
        self.send_buf = {}
        self.recv_buf = {}
        self._stream = None
        self._native = None      # True once the set moves the halos itself (init_native_comm)

    # ---- DataBus::create_custom_types ------------------------------------------------------
    def create_custom_types(self):
        req = {}
        for p in range(self.world):
            if p == self.rank:
                continue
            n_recv, _ = self.ms.halo_counts(p)
            if n_recv:
                req[p] = self.ms.halo_get_requests(p)
        allreq = [None] * self.world
        dist.all_gather_object(allreq, req, group=self.group)
        for p in range(self.world):
            if p == self.rank:
                continue
            wanted = allreq[p].get(self.rank)
            if wanted is not None and len(wanted):
                self.ms.halo_set_sends(p, wanted)
        for p in range(self.world):
            if p == self.rank:
                continue
            nr, ns = self.ms.halo_counts(p)
            if nr or ns:
                self.peers.append(p)
                self.n_recv[p], self.n_send[p] = nr, ns

    # ---- first DataBus::sync_nodes (host memory, before pre_process_meshes) ----------------
    def sync_nodes_startup(self):
        out = {p: self.ms.halo_pack_host(p) for p in self.peers if self.n_send[p]}
        allout = [None] * self.world
        dist.all_gather_object(allout, out, group=self.group)
        for p in self.peers:
            if self.n_recv[p]:
                self.ms.halo_unpack_host(p, allout[p][self.rank])

    # ---- per-stage device exchange ----------------------------------------------------------
    def _ensure_buffers(self):
        if self._stream is not None:
            return
        self._stream = torch.cuda.ExternalStream(self.ms.stream())
        for p in self.peers:
            if self.n_send[p]:
                self.send_buf[p] = torch.empty(9 * self.n_send[p], dtype=torch.float32, device="cuda")
            if self.n_recv[p]:
                self.recv_buf[p] = torch.empty(9 * self.n_recv[p], dtype=torch.float32, device="cuda")

    def sync_nodes(self):
        """DataBus::sync_nodes: same-process halo copy + grouped send/recv with every neighbour."""
        self._ensure_buffers()
        self.ms.sync_nodes()
        if not self.peers:
            return
        with torch.cuda.stream(self._stream):
            ops = []
            for p in self.peers:
                if p in self.send_buf:
                    self.ms.halo_pack(p, self.send_buf[p].data_ptr())
                    ops.append(dist.P2POp(dist.isend, self.send_buf[p], p, group=self.group))
                if p in self.recv_buf:
                    ops.append(dist.P2POp(dist.irecv, self.recv_buf[p], p, group=self.group))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for p in self.peers:
                if p in self.recv_buf:
                    self.ms.halo_unpack(p, self.recv_buf[p].data_ptr())

    # ---- library-driven exchange: the set owns an NCCL communicator --------------------------
    def init_native_comm(self):
        """Collective, after pre_process_meshes: hand the set its own NCCL communicator so that a whole
        step (halo exchange included) is ONE library call -- no per-stage Python on the critical path."""
        ids = [self.ms.comm_get_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(ids, src=0, group=self.group)
        self.ms.comm_init(ids[0], self.rank, self.world)
        self._native = True

    def do_next_step(self, tau=0.002):
        ms = self.ms
        if self._native is None:
            import os
            self._native = False
            if dist.get_backend(self.group) == "nccl" and os.environ.get("GCM_B200_TORCH_HALO", "0") != "1":
                self.init_native_comm()
        if self._native and tau == 0.002:
            ms.do_next_step()
            return
        ms.begin_step()
        for stage in range(4):
            self.sync_nodes()
            ms.do_next_part_step(tau, stage)
        ms.end_step()
