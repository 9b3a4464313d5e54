"""Slab-sharded stepping across the GPUs of one box: one process per GPU,
torch.distributed for the plumbing, NCCL send/recv of the slab edge rows / planes per RK stage.

The grid is cut along its slowest axis (y in 2-D, z in 3-D).  Every rank builds the model for its
own rows plus two halo rows of each neighbour (masks only), computes only the rows it owns, and after
each stage sends its first / last owned row of k_m (and of the new u after the last stage) to the
neighbours.  The edge rows are computed first (phase 1) so that their transfer overlaps the
interior kernel (phase 2).  There is no reduction inside the step, so the result is bit-identical to
the single-GPU run.

The exchange itself (`exchange_rows`) works on any torch tensors, which is what the world-size-2
gloo tests exercise on CPU.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import FieldView, SpatialSimulator, load_library


def partition(n, world):
    """Rows [lo, hi) owned by each rank (contiguous, sizes differ by at most one)."""
    return [((r * n) // world, ((r + 1) * n) // world) for r in range(world)]


def exchange_rows(tensors, unit, base, own_lo, own_hi, rank, world, group=None):
    """Fill the halo row below / above the owned rows of every flat tensor in `tensors` with the
    neighbours' edge rows.  Row r of a tensor is tensor[base + r*unit : base + (r+1)*unit].
    Returns the list of outstanding requests (wait on them before the halos are read)."""
    ops = []
    def row(t, r):
        return t[base + r * unit: base + (r + 1) * unit]
    for t in tensors:
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, row(t, own_lo), rank - 1, group))
            ops.append(dist.P2POp(dist.irecv, row(t, own_lo - 1), rank - 1, group))
        if rank < world - 1:
            ops.append(dist.P2POp(dist.isend, row(t, own_hi - 1), rank + 1, group))
            ops.append(dist.P2POp(dist.irecv, row(t, own_hi), rank + 1, group))
    if not ops:
        return []
    return dist.batch_isend_irecv(ops)


class _DevArray(object):
    """Zero-copy torch view of a device allocation owned by the library."""
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}


class ShardedSimulator(object):
    """The same model stepped by `world` ranks, each owning a slab."""

    def __init__(self, sbml, xdim, ydim, zdim=1, dim=2, device=0, group=None):
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.dim = dim
        n = zdim if dim == 3 else ydim
        if self.world > n:
            raise ValueError("more ranks than rows along the slab axis")
        self.lo, self.hi = partition(n, self.world)[self.rank]
        self.lib = load_library()
        self.compute = torch.cuda.current_stream()
        self.comm = torch.cuda.Stream()
        self.edge = torch.cuda.Stream(priority=-1)  # the slab's edge rows run beside the interior kernel
        self.sim = SpatialSimulator(device=device, slab=(self.lo, self.hi) if self.world > 1 else None,
                                    stream=self.compute.cuda_stream)
        self.sim.initFromString(sbml, xdim, ydim, zdim)
        self.h = C.c_void_p(self.sim.getDeviceHandle())
        self.S = self.sim.getNumStagedSpecies()
        nz, ny, nx = self.sim.localShape()
        self.held = nz if dim == 3 else ny
        # halo rows held below the owned range (2 unless the slab touches the grid edge)
        self.held_lo = max(0, self.lo - 2) if self.world > 1 else 0
        self.own_lo = self.lo - self.held_lo
        self.own_hi = self.hi - self.held_lo
        self.views = {}
        if self.world > 1:
            for s in range(self.S):
                for which in (0, 1, 2, 3, 5):
                    v = FieldView()
                    if self.lib.sb2_species_view(self.h, s, which, C.byref(v)) != 0:
                        raise RuntimeError(self.lib.sb2_last_error(self.h).decode())
                    self.views[(s, which)] = torch.as_tensor(_DevArray(v.ptr, v.n_alloc), device="cuda")
                    if dim == 3:
                        self.unit, self.base = v.plane_pitch, v.first_cell - v.row_pitch
                    else:
                        self.unit, self.base = v.row_pitch, v.first_cell
            # view 0 / 5 name "current" / "other" at creation time; track which buffer is current
            self.cur = 0
        self.time_slot = self.sim.getTimeSlot()  # sim_time = stepIndex * dt reaches the kinetic laws through this constant
        self.halo_ready = None

    def _u(self, s, current):
        which = 0 if (self.cur == 0) == current else 5
        return self.views[(s, which)]

    def _exchange(self, tensors, after_event):
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(after_event)
            reqs = exchange_rows(tensors, self.unit, self.base, self.own_lo, self.own_hi, self.rank, self.world, self.group)
            for r in reqs:
                r.wait()
        done = torch.cuda.Event()
        done.record(self.comm)
        return done

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.sb2_last_error(self.h).decode())

    def step(self, step_index, dt):
        """One RK4 step with the reference's run() convention oneStep(stepIndex, dt)."""
        if self.world == 1:
            self.sim.oneStep(step_index, dt)
            return
        lib, h = self.lib, self.h
        self._ck(lib.sb2_begin_step(h, float(step_index), float(dt), self.time_slot))
        fused = lib.sb2_combine_is_fused(h) == 1
        # Per stage: the two edge rows on their own stream (they need the neighbours' halo rows of the previous stage),
        # the interior rows on the compute stream (they need only this rank's rows), the halo exchange of the new edge
        # rows on the communication stream as soon as the edge kernel is done — all three overlap.
        pending = self.halo_ready      # halos of u from the previous step (None at the first step)
        prev_interior = None
        prev_edge = None
        for m in range(4):
            if pending is not None:
                self.edge.wait_event(pending)
            if prev_interior is not None:
                self.edge.wait_event(prev_interior)
            else:
                self.edge.wait_stream(self.compute)
            if prev_edge is not None:
                self.compute.wait_event(prev_edge)
            self._ck(lib.sb2_stage_on(h, m, 1, C.c_void_p(self.edge.cuda_stream)))
            ev = torch.cuda.Event()
            ev.record(self.edge)
            prev_edge = ev
            if m < 3:
                pending = self._exchange([self.views[(s, 1 + m)] for s in range(self.S)], ev)
            elif fused:
                pending = self._exchange([self._u(s, current=False) for s in range(self.S)], ev)
            self._ck(lib.sb2_stage(h, m, 2))      # interior rows
            prev_interior = torch.cuda.Event()
            prev_interior.record(self.compute)
        self.compute.wait_event(prev_edge)
        self._ck(lib.sb2_end_step(h))
        self.sim.deviceStateChanged()  # the host mirrors of the species are stale now
        if fused:
            self.cur ^= 1
        else:
            ev = torch.cuda.Event()
            ev.record(self.compute)
            pending = self._exchange([self._u(s, current=True) for s in range(self.S)], ev)
        self.halo_ready = pending
        if pending is not None:
            self.compute.wait_event(pending)

    def run_steps(self, first_step_index, dt, nsteps):
        if self.world == 1:
            self.sim.runSteps(first_step_index, dt, nsteps)
            return
        for i in range(nsteps):
            self.step(first_step_index + i, dt)

    def owned(self, species):
        """Compact values of the rows this rank owns (numpy)."""
        a = self.sim.getVariableCompact(species)
        if self.world == 1:
            return a
        return a[self.own_lo:self.own_hi] if self.dim == 3 else a[:, self.own_lo:self.own_hi]

    def num_domain_cells(self):
        return self.sim.getNumDomainCells()

    def launch_count(self):
        return self.sim.getDeviceLaunchCount()
