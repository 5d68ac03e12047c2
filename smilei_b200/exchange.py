"""Neighbour exchange between patches held by different GPUs: the NCCL replacement of
SmileiMPI / SyncVectorPatch for the hot path.

One patch per process per GPU on a Cartesian grid of ranks (`npatch`), periodic box.  Three
exchanges per step, each x -> y -> z sequential so that edges and corners propagate exactly
as in the reference:

  * sum_J       SyncVectorPatch::sumAllComponents          (src/Patch/SyncVectorPatch.cpp:203-…)
  * exchange_B  SyncVectorPatch::exchangeB                  (:667-699, :1441-1527)
  * particles   SyncVectorPatch::initExchParticles / finalizeExchParticlesAndSort (:27-113),
                Patch::exchNbrOfParticles / prepareParticles / exchParticles / cornersParticles
                (src/Patch/Patch.cpp:560-800)

Transport is torch.distributed point-to-point (`batch_isend_irecv`): NCCL over NVLink on
GPUs, gloo in the CPU tests.  NCCL has no tags; when the -dim and +dim neighbour are the
same peer (2 ranks along a periodic dimension) both payloads travel in ONE message with a
fixed layout, which replaces the MPI tags of Patch.cpp:573,587.  A dimension held by a
single rank wraps on the device without any message (the reference's same-rank pointer
copies, SyncVectorPatch.cpp:288-311).

The patch object only has to provide the halo / migration hooks of include/smilei_b200.h
(smilei_b200.capi.Patch does; the CPU tests plug an oracle-backed stand-in).
"""
import numpy as np
import torch
import torch.distributed as dist

UNPACK_COPY, UNPACK_ADD = 0, 1
RECORD = 8  # doubles per migrating particle (SB200_PARTICLE_RECORD_DOUBLES)

J_FIELDS = ("Jx", "Jy", "Jz")
# components exchanged per direction = the two that are dual in it (SyncVectorPatch.cpp:690-695)
B_EXCHANGE = ((0, ("By", "Bz")), (1, ("Bx", "Bz")), (2, ("Bx", "By")))
_DUAL = {"Jx": (1, 0, 0), "Jy": (0, 1, 0), "Jz": (0, 0, 1), "Bx": (0, 1, 1), "By": (1, 0, 1), "Bz": (1, 1, 0),
         "rho": (0, 0, 0)}


def rank_to_pcoord(rank, npatch):
    """Row-major rank <-> patch coordinates (x slowest).  Inside one NVSwitch box every peer is
    equidistant, so the Hilbert ordering of the reference brings nothing."""
    return tuple(int(v) for v in np.unravel_index(rank, npatch))


def pcoord_to_rank(pcoord, npatch):
    return int(np.ravel_multi_index(tuple(c % n for c, n in zip(pcoord, npatch)), npatch))


class Exchanger:
    """Halo and particle exchange for ONE patch (this rank's) on a periodic Cartesian rank grid."""

    def __init__(self, patch, n, oversize, cell_length, npatch, pcoord, device, group=None, particle_buffer=1 << 16,
                 periodic=(True, True, True)):
        self.patch = patch
        self.n = tuple(n)
        self.o = tuple(oversize)
        self.cell = tuple(cell_length)
        self.npatch = tuple(npatch)
        self.pcoord = tuple(pcoord)
        self.device = torch.device(device)
        self.group = group
        self.world = int(np.prod(self.npatch))
        if self.world > 1:
            assert dist.is_initialized(), "torch.distributed must be initialised for more than one patch"
            assert dist.get_world_size(group) == self.world
            assert pcoord_to_rank(self.pcoord, self.npatch) == dist.get_rank(group)
        # a non-periodic dimension has no neighbour beyond the box (MPI_PROC_NULL in Patch::neighbor_): None
        self.periodic = tuple(bool(v) for v in periodic)
        self.nbr = []
        for d in range(3):
            lo = list(self.pcoord)
            hi = list(self.pcoord)
            lo[d] -= 1
            hi[d] += 1
            r_lo, r_hi = pcoord_to_rank(lo, self.npatch), pcoord_to_rank(hi, self.npatch)
            if not self.periodic[d]:
                if self.pcoord[d] == 0:
                    r_lo = None
                if self.pcoord[d] == self.npatch[d] - 1:
                    r_hi = None
            self.nbr.append((r_lo, r_hi))
        self._bufs = {}
        self._pcap = int(particle_buffer)
        self.bytes_sent = 0

    # ------------------------------------------------------------------ helpers
    def _buf(self, key, numel):
        b = self._bufs.get(key)
        if b is None or b.numel() < numel:
            b = torch.empty(max(numel, 1), dtype=torch.float64, device=self.device)
            self._bufs[key] = b
        return b

    def _sendrecv(self, ops):
        """ops: list of (peer, send_tensor, recv_tensor).  One batch = one NCCL group."""
        p2p = []
        for peer, s, r in ops:
            p2p.append(dist.P2POp(dist.isend, s, peer, self.group))
            p2p.append(dist.P2POp(dist.irecv, r, peer, self.group))
            self.bytes_sent += s.numel() * s.element_size()
        for w in dist.batch_isend_irecv(p2p):
            w.wait()

    def _exchange_slabs(self, dim, to_lo, to_hi, size_from_lo=None, size_from_hi=None, tag=""):
        """Send `to_lo` to the -dim neighbour and `to_hi` to the +dim one; returns (from_lo, from_hi) = what
        those neighbours sent to me.  By default the exchange is symmetric (I receive from a side as much as
        I send to it); for particle records the incoming sizes are given.  `tag` names the caller's own set of
        staging buffers: the particle exchange runs on a side stream concurrently with the J sum and the B exchange
        (Simulation.exchange_and_solve), so the three must not share send / receive buffers."""
        lo, hi = self.nbr[dim]
        n_lo, n_hi = to_lo.numel(), to_hi.numel()
        m_lo = n_hi if size_from_lo is None else size_from_lo     # the -dim neighbour sends me ITS to_hi block
        m_hi = n_lo if size_from_hi is None else size_from_hi
        if lo is None or hi is None:
            # box side without neighbour: nothing goes or comes that way (callers skip that side's unpack)
            from_lo = self._buf((tag + "rl", dim), m_lo)[:0 if lo is None else m_lo]
            from_hi = self._buf((tag + "rh", dim), m_hi)[:0 if hi is None else m_hi]
            ops = []
            if lo is not None:
                ops.append((lo, to_lo, from_lo))
            if hi is not None:
                ops.append((hi, to_hi, from_hi))
            if ops:
                self._sendrecv(ops)
            return from_lo, from_hi
        if lo == hi:
            # one peer on both sides: [payload for its +side | payload for its -side] in one message
            send = self._buf((tag + "s2", dim), n_lo + n_hi)[:n_lo + n_hi]
            send[:n_lo].copy_(to_lo)
            send[n_lo:].copy_(to_hi)
            recv = self._buf((tag + "r2", dim), m_hi + m_lo)[:m_hi + m_lo]
            self._sendrecv([(lo, send, recv)])
            # the peer's first block was meant for its -dim neighbour's +side = my +side
            return recv[m_hi:], recv[:m_hi]
        from_lo = self._buf((tag + "rl", dim), m_lo)[:m_lo]
        from_hi = self._buf((tag + "rh", dim), m_hi)[:m_hi]
        self._sendrecv([(lo, to_lo, from_lo), (hi, to_hi, from_hi)])
        return from_lo, from_hi

    # ------------------------------------------------------------------ J
    def sum_J(self, fields=J_FIELDS):
        """Both sides of every shared plane end up holding the sum (SyncVectorPatch.cpp:296-311)."""
        p = self.patch
        for dim in range(3):
            if self.npatch[dim] == 1:
                if self.periodic[dim]:
                    for f in fields:
                        p.halo_sum_self(f, dim)
                continue
            has_lo, has_hi = self.nbr[dim][0] is not None, self.nbr[dim][1] is not None
            sizes, gsp = [], []
            for f in fields:
                g = 1 + 2 * self.o[dim] + _DUAL[f[0] if isinstance(f, tuple) else f][dim]   # SyncVectorPatch.cpp:235-237,284; (name, ispec): a species' own array
                gsp.append(g)
                sizes.append(g * p.halo_plane_elems(f, dim))
            tot = sum(sizes)
            to_lo = self._buf(("jl", dim), tot)[:tot]
            to_hi = self._buf(("jh", dim), tot)[:tot]
            off = 0
            for f, g, s in zip(fields, gsp, sizes):
                p.halo_pack(f, dim, 0, g, to_lo[off:off + s].data_ptr())            # my planes [0,gsp)
                p.halo_pack(f, dim, self.n[dim], g, to_hi[off:off + s].data_ptr())  # my planes [n,n+gsp)
                off += s
            self._sync_patch_stream()
            from_lo, from_hi = self._exchange_slabs(dim, to_lo, to_hi, tag="J")
            off = 0
            for f, g, s in zip(fields, gsp, sizes):
                # the -dim neighbour's [n,n+gsp) planes are my [0,gsp); the +dim one's [0,gsp) are my [n,n+gsp)
                if has_lo:
                    p.halo_unpack(f, dim, 0, g, from_lo[off:off + s].data_ptr(), UNPACK_ADD)
                if has_hi:
                    p.halo_unpack(f, dim, self.n[dim], g, from_hi[off:off + s].data_ptr(), UNPACK_ADD)
                off += s

    # ------------------------------------------------------------------ B
    def exchange_B(self):
        """R[0,o) <- L[n,n+o) ; L[n+gsp,n+gsp+o) <- R[gsp,gsp+o), gsp = o+2 (SyncVectorPatch.cpp:1501-1527)."""
        p = self.patch
        for dim, comps in B_EXCHANGE:
            o = self.o[dim]
            gsp = o + 2
            if self.npatch[dim] == 1:
                if self.periodic[dim]:
                    for f in comps:
                        p.halo_exchange_self(f, dim)
                continue
            has_lo, has_hi = self.nbr[dim][0] is not None, self.nbr[dim][1] is not None
            sizes = [o * p.halo_plane_elems(f, dim) for f in comps]
            tot = sum(sizes)
            to_lo = self._buf(("bl", dim), tot)[:tot]
            to_hi = self._buf(("bh", dim), tot)[:tot]
            off = 0
            for f, s in zip(comps, sizes):
                p.halo_pack(f, dim, gsp, o, to_lo[off:off + s].data_ptr())           # I am R of my -dim neighbour
                p.halo_pack(f, dim, self.n[dim], o, to_hi[off:off + s].data_ptr())   # I am L of my +dim neighbour
                off += s
            self._sync_patch_stream()
            from_lo, from_hi = self._exchange_slabs(dim, to_lo, to_hi)
            off = 0
            for f, s in zip(comps, sizes):
                if has_lo:
                    p.halo_unpack(f, dim, 0, o, from_lo[off:off + s].data_ptr(), UNPACK_COPY)
                if has_hi:
                    p.halo_unpack(f, dim, self.n[dim] + gsp, o, from_hi[off:off + s].data_ptr(), UNPACK_COPY)
                off += s

    # ------------------------------------------------------------------ particles
    def exchange_particles(self, n_species):
        """x, then y, then z; arrivals are re-tagged on unpack so corner particles are forwarded in
        the next dimension (Patch::cornersParticles).  All species travel together: per dimension one
        message with the counts and one with the records per neighbour."""
        p = self.patch
        pack = getattr(p, "leaving_pack_known", None)
        for dim in range(3):
            L = self.cell[dim] * float(self.n[dim] * self.npatch[dim])      # Patch.cpp:626
            wrap_lo = L if self.pcoord[dim] == 0 else 0.                     # Patch.cpp:636-642
            wrap_hi = -L if self.pcoord[dim] == self.npatch[dim] - 1 else 0.  # Patch.cpp:643-649
            counts = [p.leaving_count(s) for s in range(n_species)]          # Patch::exchNbrOfParticles: sizes first
            c_lo = [c[2 * dim] for c in counts]
            c_hi = [c[2 * dim + 1] for c in counts]
            if self.npatch[dim] == 1 and not self.periodic[dim]:
                # nobody to exchange with: the particle boundary condition (remove) has dealt with the leavers
                assert not any(c_lo) and not any(c_hi), "particles tagged for exchange across a non-periodic box side"
                continue
            if self.npatch[dim] == 1:
                for s in range(n_species):
                    self._ensure_pcap(max(c_lo[s], c_hi[s]))
                    to_lo = self._buf(("pl", dim), RECORD * self._pcap)
                    to_hi = self._buf(("ph", dim), RECORD * self._pcap)
                    pack(s, dim, 0, wrap_lo, to_lo.data_ptr(), self._pcap, c_lo[s])
                    pack(s, dim, 1, wrap_hi, to_hi.data_ptr(), self._pcap, c_hi[s])
                    p.arriving_unpack(s, to_lo.data_ptr(), c_lo[s])
                    p.arriving_unpack(s, to_hi.data_ptr(), c_hi[s])
                continue
            has_lo, has_hi = self.nbr[dim][0] is not None, self.nbr[dim][1] is not None
            cnt_lo = torch.tensor([float(v) for v in c_lo], dtype=torch.float64, device=self.device)
            cnt_hi = torch.tensor([float(v) for v in c_hi], dtype=torch.float64, device=self.device)
            r_lo, r_hi = self._exchange_slabs(dim, cnt_lo, cnt_hi, tag="C")
            n_from_lo = [int(v) for v in r_lo.cpu().tolist()] if has_lo else [0] * n_species
            n_from_hi = [int(v) for v in r_hi.cpu().tolist()] if has_hi else [0] * n_species
            # per direction, every species block is padded to the largest count of that direction, which both
            # ends of the link know: what I send to -dim is what that neighbour announced-to-receive, etc.
            pad_to_lo, pad_to_hi = max(max(c_lo), 1), max(max(c_hi), 1)
            pad_from_lo, pad_from_hi = max(max(n_from_lo), 1), max(max(n_from_hi), 1)
            to_lo = self._buf(("pl", dim), RECORD * pad_to_lo * n_species)[:RECORD * pad_to_lo * n_species]
            to_hi = self._buf(("ph", dim), RECORD * pad_to_hi * n_species)[:RECORD * pad_to_hi * n_species]
            for s in range(n_species):
                pack(s, dim, 0, wrap_lo, to_lo[RECORD * pad_to_lo * s:].data_ptr(), pad_to_lo, c_lo[s])
                pack(s, dim, 1, wrap_hi, to_hi[RECORD * pad_to_hi * s:].data_ptr(), pad_to_hi, c_hi[s])
            self._sync_patch_stream()
            from_lo, from_hi = self._exchange_slabs(dim, to_lo, to_hi, RECORD * pad_from_lo * n_species,
                                                    RECORD * pad_from_hi * n_species, tag="P")
            for s in range(n_species):
                # arrivals from the -dim neighbour first, then from the +dim one (deterministic order)
                if has_lo:
                    p.arriving_unpack(s, from_lo[RECORD * pad_from_lo * s:].data_ptr(), n_from_lo[s])
                if has_hi:
                    p.arriving_unpack(s, from_hi[RECORD * pad_from_hi * s:].data_ptr(), n_from_hi[s])
            self._sync_patch_stream()

    # ------------------------------------------------------------------ moving window across ranks along x
    WINDOW_FIELDS = ("Ex", "Ey", "Ez", "Bx", "By", "Bz", "Bxm", "Bym", "Bzm")
    _DUAL_X = {"Ex": 1, "Ey": 0, "Ez": 0, "Bx": 0, "By": 1, "Bz": 1, "Bxm": 0, "Bym": 1, "Bzm": 1}

    def _send_left_recv_right(self, send, recv):
        """One-directional hand-over along x: my block goes to the -x neighbour, the +x neighbour's comes to me."""
        lo, hi = self.nbr[0]
        p2p = []
        if lo is not None and send is not None:
            p2p.append(dist.P2POp(dist.isend, send, lo, self.group))
            self.bytes_sent += send.numel() * send.element_size()
        if hi is not None and recv is not None:
            p2p.append(dist.P2POp(dist.irecv, recv, hi, self.group))
        if p2p:
            for w in dist.batch_isend_irecv(p2p):
                w.wait()

    def window_fields_begin(self, stride):
        """Before the shift: the planes the -x neighbour will uncover at its right end are my (old) planes
        [2*oversize+1+dual, +stride) of E, B, B_m; returns what the +x neighbour sent (or None at the right end)."""
        if self.npatch[0] == 1:
            return None
        p = self.patch
        lo, hi = self.nbr[0]
        sizes = [stride * p.halo_plane_elems(f, 0) for f in self.WINDOW_FIELDS]
        tot = sum(sizes)
        send = recv = None
        if lo is not None:
            send = self._buf(("wl", 0), tot)[:tot]
            off = 0
            for f, s in zip(self.WINDOW_FIELDS, sizes):
                p.halo_pack(f, 0, 2 * self.o[0] + 1 + self._DUAL_X[f], stride, send[off:off + s].data_ptr())
                off += s
            self._sync_patch_stream()
        if hi is not None:
            recv = self._buf(("wr", 0), tot)[:tot]
        self._send_left_recv_right(send, recv)
        return recv

    def window_fields_end(self, stride, recv):
        """After the shift: the +x neighbour's planes become my last `stride` real planes."""
        if recv is None:
            return
        p = self.patch
        off = 0
        for f in self.WINDOW_FIELDS:
            s = stride * p.halo_plane_elems(f, 0)
            nreal = self.n[0] + 2 * self.o[0] + 1 + self._DUAL_X[f]
            p.halo_unpack(f, 0, nreal - stride, stride, recv[off:off + s].data_ptr(), UNPACK_COPY)
            off += s
        self._sync_patch_stream()

    def exchange_window_particles(self, n_species):
        """Particles the shift left behind on a patch with a -x neighbour (tag -2) move to that neighbour."""
        if self.npatch[0] == 1:
            return
        p = self.patch
        lo, hi = self.nbr[0]
        counts = [p.leaving_count(s)[0] for s in range(n_species)]
        cnt = torch.tensor([float(v) for v in counts], dtype=torch.float64, device=self.device)
        rcnt = torch.zeros(n_species, dtype=torch.float64, device=self.device)
        self._send_left_recv_right(cnt if lo is not None else None, rcnt if hi is not None else None)
        n_from = [int(v) for v in rcnt.cpu().tolist()] if hi is not None else [0] * n_species
        pad_to, pad_from = max(max(counts), 1), max(max(n_from), 1)
        send = recv = None
        if lo is not None:
            send = self._buf(("pwl", 0), RECORD * pad_to * n_species)[:RECORD * pad_to * n_species]
            for s in range(n_species):
                got = p.leaving_pack(s, 0, 0, 0., send[RECORD * pad_to * s:].data_ptr(), pad_to)
                assert got == counts[s]
            self._sync_patch_stream()
        if hi is not None:
            recv = self._buf(("pwr", 0), RECORD * pad_from * n_species)[:RECORD * pad_from * n_species]
        self._send_left_recv_right(send, recv)
        if hi is not None:
            for s in range(n_species):
                p.arriving_unpack(s, recv[RECORD * pad_from * s:].data_ptr(), n_from[s])
            self._sync_patch_stream()

    def _ensure_pcap(self, need):
        if need > self._pcap:
            self._pcap = int(need * 1.5) + 16

    def _sync_patch_stream(self):
        # Pack / unpack kernels run on the patch's stream and the collective on torch's current stream.  On the GPU
        # the two are kept THE SAME stream: Simulation hands torch's current stream to the library at start-up
        # (sb200_patch_set_stream) and this check refuses to go on if the caller switched torch's stream since —
        # NCCL would otherwise order against a stream the pack kernels do not run on.  The CPU stand-in of the
        # tests synchronises here.
        if self.device.type == "cuda" and not getattr(self.patch, "own_stream", False):
            import torch
            bound = getattr(self.patch, "bound_stream", None)
            if bound is not None and torch.cuda.current_stream(self.device).cuda_stream != bound:
                raise RuntimeError("torch's current CUDA stream changed after the Simulation was created: the exchange "
                                   "kernels and the collectives would no longer be stream-ordered")
            return
        sync = getattr(self.patch, "synchronize", None)
        if sync is not None:
            sync()
