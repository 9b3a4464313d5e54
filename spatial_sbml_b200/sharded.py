"""Slab-sharded stepping across the GPUs of one box: one process per GPU,
torch.distributed for the plumbing, NCCL send/recv of the slab edge rows / planes.

The grid is cut along its slowest axis (y in 2-D, z in 3-D).  Every rank builds the model for its
own rows plus two halo rows of each neighbour and computes only the rows it owns.  The step itself is
orchestrated INSIDE the library (sb2_sharded_begin / sb2_sharded_next, include/spatial_b200.h): it
runs the edge rows first on a high-priority stream, hands back the list of rows that have to travel
(k_m after each stage, the new u after the last one, u and the face values after every advection
pass, ...) together with the stream to queue the transfer on, and runs the interior rows while the
transfer is in flight.  This module only moves the rows it is handed.  There is no reduction inside
the step, so the result is bit-identical to the single-GPU run.

The exchange itself (`exchange_rows`) works on any torch tensors, which is what the world-size-2
gloo tests exercise on CPU.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import SpatialSimulator, load_library

SB2_MAX_XBUF = 40


class Exchange(C.Structure):
    """sb2_exchange of include/spatial_b200.h"""
    _fields_ = [("n", C.c_int32), ("ptr", C.c_void_p * SB2_MAX_XBUF), ("width", C.c_int32), ("unit", C.c_int64), ("base", C.c_int64),
                ("own_lo", C.c_int32), ("own_hi", C.c_int32), ("has_lo", C.c_int32), ("has_hi", C.c_int32),
                ("stream", C.c_void_p), ("ready_event", C.c_void_p), ("device", C.c_int32)]


def partition(n, world):
    """Rows [lo, hi) owned by each rank (contiguous, sizes differ by at most one)."""
    return [((r * n) // world, ((r + 1) * n) // world) for r in range(world)]


def exchange_rows(tensors, unit, base, own_lo, own_hi, rank, world, group=None, width=1):
    """Fill the `width` halo rows below / above the owned rows of every flat tensor in `tensors` with the
    neighbours' edge rows.  Row r of a tensor is tensor[base + r*unit : base + (r+1)*unit].
    Returns the list of outstanding requests (wait on them before the halos are read)."""
    ops = []
    def rows(t, r):
        return t[base + r * unit: base + (r + width) * unit]
    for t in tensors:
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, rows(t, own_lo), rank - 1, group))
            ops.append(dist.P2POp(dist.irecv, rows(t, own_lo - width), rank - 1, group))
        if rank < world - 1:
            ops.append(dist.P2POp(dist.isend, rows(t, own_hi - width), rank + 1, group))
            ops.append(dist.P2POp(dist.irecv, rows(t, own_hi), rank + 1, group))
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
        self.sim = SpatialSimulator(device=device, slab=(self.lo, self.hi) if self.world > 1 else None,
                                    stream=self.compute.cuda_stream)
        self.sim.initFromString(sbml, xdim, ydim, zdim)
        self.h = C.c_void_p(self.sim.getDeviceHandle())
        self.S = self.sim.getNumStagedSpecies()
        nz, ny, nx = self.sim.localShape()
        self.held = nz if dim == 3 else ny
        # first row held: the owned range plus the halo rows of the neighbouring slab below it
        self.held_lo = self.sim.getSlabRows()[0] if self.world > 1 else 0
        self.own_lo = self.lo - self.held_lo
        self.own_hi = self.hi - self.held_lo
        self.time_slot = self.sim.getTimeSlot()  # sim_time = stepIndex * dt reaches the kinetic laws through this constant
        # every rank knows every slab's size (they differ by at most one row): with eight rows or more per slab the library
        # exchanges four rows of u once per step and recomputes the apron rows, instead of exchanging after every stage
        # (models off the fused schedule keep the per-stage exchanges); SB2_NO_APRON=1 keeps them for every model
        # (2-D grids: in 3-D the apron is twelve whole planes per slab and step, measured slower than the exchanges it saves)
        self.apron = self.world > 1 and n // self.world >= 8 and not os.environ.get("SB2_NO_APRON") and (dim == 2 or bool(os.environ.get("SB2_APRON")))
        if self.apron:
            self.lib.sb2_sharded_set_apron(self.h, 1)
        self._x = Exchange()
        self._streams = {}
        self._views = {}
        self.exchanges = 0

    def _ck(self, rc):
        if rc < 0:
            raise RuntimeError(self.lib.sb2_last_error(self.h).decode())
        return rc

    def _rows(self, ptr, first_row, x):
        """torch view of rows [first_row, first_row + width) of one buffer of an exchange"""
        off = x.base + first_row * x.unit
        key = (ptr, off, x.width * x.unit)
        v = self._views.get(key)
        if v is None:
            v = torch.as_tensor(_DevArray(ptr + 8 * off, x.width * x.unit), device="cuda")
            self._views[key] = v
        return v

    def _exchange(self, x):
        """Move the rows the library asked for, ordered on the stream it named."""
        st = self._streams.get(x.stream)
        if st is None:
            st = torch.cuda.ExternalStream(x.stream)
            self._streams[x.stream] = st
        ops = []
        w = x.width
        for b in range(x.n):
            p = x.ptr[b]
            if x.has_lo:
                ops.append(dist.P2POp(dist.isend, self._rows(p, x.own_lo, x), self.rank - 1, self.group))
                ops.append(dist.P2POp(dist.irecv, self._rows(p, x.own_lo - w, x), self.rank - 1, self.group))
            if x.has_hi:
                ops.append(dist.P2POp(dist.isend, self._rows(p, x.own_hi - w, x), self.rank + 1, self.group))
                ops.append(dist.P2POp(dist.irecv, self._rows(p, x.own_hi, x), self.rank + 1, self.group))
        if not ops:
            return
        with torch.cuda.stream(st):
            for r in dist.batch_isend_irecv(ops):
                r.wait()  # stream-ordered: the library's stream now waits for the transfer
        self.exchanges += 1

    def step(self, step_index, dt):
        """One RK4 step with the reference's run() convention oneStep(stepIndex, dt)."""
        if self.world == 1:
            self.sim.oneStep(step_index, dt)
            return
        lib, h, x = self.lib, self.h, self._x
        self._ck(lib.sb2_sharded_begin(h, float(step_index), float(dt), self.time_slot))
        while self._ck(lib.sb2_sharded_next(h, C.byref(x))) == 1:
            self._exchange(x)
        self.sim.deviceStateChanged()  # the host mirrors of the species are stale now

    def step_host(self, step_index, dt, names, inputs, outputs, nslabs=0):
        """One step whose species values come from / go to HOST arrays (one per name, shaped like sim.localShape(), ideally pinned):
        upload, compute and download of this rank's slab are pipelined over sub-slabs; the edge rows of u travel to the neighbouring
        ranks once per step and the apron rows are recomputed (sb2_step_host_begin / _finish)."""
        if self.world == 1:
            self.sim.oneStepHost(step_index, dt, dict(zip(names, inputs)), dict(zip(names, outputs)), nslabs)
            return
        dp = C.POINTER(C.c_double)
        n = len(names)
        # species ids on the device follow the model's species order; names are given in that order by the caller
        cin = (dp * self.S)(*[inputs[i].ctypes.data_as(dp) if i < n else dp() for i in range(self.S)])
        cout = (dp * self.S)(*[outputs[i].ctypes.data_as(dp) if i < n else dp() for i in range(self.S)])
        x = self._x
        rc = self._ck(self.lib.sb2_step_host_begin(self.h, float(step_index), float(dt), self.time_slot, cin, int(nslabs), C.byref(x)))
        if rc == 1:
            self._exchange(x)
        self._ck(self.lib.sb2_step_host_finish(self.h, cout))
        self.sim.deviceStateChanged()

    def run_steps(self, first_step_index, dt, nsteps):
        if self.world == 1:
            self.sim.runSteps(first_step_index, dt, nsteps)
            return
        for i in range(nsteps):
            self.step(first_step_index + i, dt)

    def stable_time_step(self, dt):
        """getStableTimeStep over the whole grid: the smallest admissible step of any slab (one 8-byte MIN all-reduce)."""
        v = self.sim.getStableTimeStep(dt)
        if self.world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return float(t[0])

    def owned(self, species):
        """Compact values of the rows this rank owns (numpy)."""
        a = self.sim.getVariableCompact(species)
        if self.world == 1:
            return a
        return a[self.own_lo:self.own_hi] if self.dim == 3 else a[:, self.own_lo:self.own_hi]

    def owned_of(self, local_array):
        """The owned rows of an array shaped like sim.localShape()"""
        if self.world == 1:
            return local_array
        return local_array[self.own_lo:self.own_hi] if self.dim == 3 else local_array[:, self.own_lo:self.own_hi]

    def num_domain_cells(self):
        return self.sim.getNumDomainCells()

    def launch_count(self):
        return self.sim.getDeviceLaunchCount()
