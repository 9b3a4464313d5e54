"""spatial_sbml_b200 — B200-native explicit reaction-diffusion-advection step of spatial-sbml.

The product is the shared library ``libspatialb200.so`` (hand-written sm_100a CUDA kernels behind
the C ABI of ``include/spatial_b200.h``, plus the host-side ``SpatialSimulator`` class with the
reference's interface and its flat C binding ``include/spatial_sim_c.h``).  This module is the
thin ctypes binding used by the tests and by ``bench.py``; it contains no numerics.

There is no CPU fallback: if the library cannot be loaded an ImportError-like RuntimeError is
raised, and stepping without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspatialb200.so")
_lib = None

# public entry points of include/spatial_b200.h and include/spatial_sim_c.h (checked by the tests
# against the headers)
SB2_OK = 0
SB2_F_DOMAIN, SB2_F_XP, SB2_F_XM, SB2_F_YP, SB2_F_YM, SB2_F_ZP, SB2_F_ZM = 1, 2, 4, 8, 16, 32, 64
SB2_TOK_OP, SB2_TOK_VAR, SB2_TOK_FIELD, SB2_TOK_CONST, SB2_TOK_NOP, SB2_TOK_MEMVAR = 0, 1, 2, 3, 4, 5
SB2_MEM_SPECIES = 0x1000
SB2_SRC_NONE, SB2_SRC_CONST, SB2_SRC_FIELD = 0, 1, 2
SB2_BC_NONE, SB2_BC_NEUMANN, SB2_BC_DIRICHLET = 0, 1, 2
OPS = dict(PLUS=1, MINUS=2, TIMES=3, DIVIDE=4, POWER=5, LOG=6, AND=7, OR=8, XOR=9, EQ=10, GEQ=11, GT=12,
           LEQ=13, LT=14, NEQ=15, ABS=32, ARCCOS=33, ARCCOSH=34, ARCCSC=35, ARCSIN=36, ARCTAN=37, COS=38,
           COSH=39, CSC=40, EXP=41, LN=42, ROOT=43, SIN=44, SINH=45, TAN=46, TANH=47, NOT=48, POP_ONLY=64)


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("dim", C.c_int32),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("device", C.c_int32),
                ("own_lo", C.c_int32), ("own_hi", C.c_int32), ("row0", C.c_int32)]


class Token(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("op", C.c_uint8), ("slot", C.c_uint16)]


class Ref(C.Structure):
    _fields_ = [("species", C.c_int32), ("is_variable", C.c_int32), ("stoichiometry", C.c_double)]


class MemPoint(C.Structure):
    _fields_ = [("axis", C.c_int32), ("abs_normal", C.c_double)]


class FieldView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("row_pitch", C.c_int64), ("plane_pitch", C.c_int64),
                ("first_cell", C.c_int64), ("n_alloc", C.c_int64)]


def _proto(lib):
    vp, ci, cd, cs = C.c_void_p, C.c_int, C.c_double, C.c_char_p
    dp = C.POINTER(C.c_double)
    P = {
        # --- include/spatial_b200.h
        "sb2_create": (ci, [C.POINTER(Grid), C.POINTER(vp)]),
        "sb2_destroy": (None, [vp]),
        "sb2_last_error": (cs, [vp]),
        "sb2_set_stream": (ci, [vp, vp]),
        "sb2_add_geometry": (ci, [vp, C.POINTER(C.c_uint8)]),
        "sb2_add_species": (ci, [vp, ci, dp]),
        "sb2_add_field": (ci, [vp, dp]),
        "sb2_set_constants": (ci, [vp, dp, ci]),
        "sb2_set_constant": (ci, [vp, ci, cd]),
        "sb2_add_reaction": (ci, [vp, ci, ci, C.POINTER(Token), ci, C.POINTER(Ref), ci, ci]),
        "sb2_set_diffusion": (ci, [vp, ci, ci, ci, ci]),
        "sb2_set_advection": (ci, [vp, ci, ci, ci, ci]),
        "sb2_set_advection_faces": (ci, [vp, ci, ci, ci]),
        "sb2_set_boundary": (ci, [vp, ci, ci, ci, ci, ci]),
        "sb2_add_mem_transport": (ci, [vp, C.POINTER(Token), ci, C.POINTER(Ref), ci, ci, C.POINTER(MemPoint), ci,
                                       C.POINTER(C.c_int32), dp, C.POINTER(C.c_int32)]),
        "sb2_add_point_set": (ci, [vp, ci]),
        "sb2_add_mem_species": (ci, [vp, ci, dp]),
        "sb2_add_mem_reaction": (ci, [vp, ci, ci, C.POINTER(Token), ci, C.POINTER(Ref), ci, ci, C.POINTER(C.c_int32), ci, dp]),
        "sb2_set_mem_diffusion": (ci, [vp, ci, ci, ci, dp, C.POINTER(C.c_int32), ci, C.POINTER(C.c_int32), C.POINTER(C.c_int32), dp, dp, dp]),
        "sb2_set_mem_update_points": (ci, [vp, ci, C.POINTER(C.c_int32), ci]),
        "sb2_set_mem_fill": (ci, [vp, ci, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), ci]),
        "sb2_download_mem_species": (ci, [vp, ci, dp]),
        "sb2_upload_mem_species": (ci, [vp, ci, dp]),
        "sb2_download_mem_stage": (ci, [vp, ci, ci, dp]),
        "sb2_add_assignment_rule": (ci, [vp, ci, ci, ci, C.POINTER(Token), ci]),
        "sb2_finalize": (ci, [vp]),
        "sb2_step": (ci, [vp, cd, cd, cd, ci, ci]),
        "sb2_step_host": (ci, [vp, cd, cd, ci, C.POINTER(dp), C.POINTER(dp), ci]),
        "sb2_step_host_begin": (ci, [vp, cd, cd, ci, C.POINTER(dp), ci, vp]),
        "sb2_step_host_finish": (ci, [vp, C.POINTER(dp)]),
        "sb2_sharded_set_apron": (ci, [vp, ci]),
        "sb2_group_step_host": (ci, [C.POINTER(vp), ci, cd, cd, ci, C.POINTER(C.POINTER(dp)), C.POINTER(C.POINTER(dp)), ci]),
        "sb2_download_species": (ci, [vp, ci, dp]),
        "sb2_upload_species": (ci, [vp, ci, dp]),
        "sb2_download_stage": (ci, [vp, ci, ci, dp]),
        "sb2_upload_field": (ci, [vp, ci, dp]),
        "sb2_download_field": (ci, [vp, ci, dp]),
        "sb2_upload_faces": (ci, [vp, ci, ci, dp]),
        "sb2_download_faces": (ci, [vp, ci, ci, dp]),
        "sb2_min_dt": (ci, [vp, cd, dp]),
        "sb2_time_advection": (ci, [vp, cd, ci, dp, C.POINTER(ci)]),
        "sb2_species_view": (ci, [vp, ci, ci, C.POINTER(FieldView)]),
        "sb2_begin_step": (ci, [vp, cd, cd, ci]),
        "sb2_stage": (ci, [vp, ci, ci]),
        "sb2_stage_on": (ci, [vp, ci, ci, vp]),
        "sb2_end_step": (ci, [vp]),
        "sb2_combine_is_fused": (ci, [vp]),
        "sb2_sharded_begin": (ci, [vp, cd, cd, ci]),
        "sb2_sharded_next": (ci, [vp, vp]),
        "sb2_group_step": (ci, [C.POINTER(vp), ci, cd, cd, ci]),
        "sb2_launch_count": (C.c_int64, [vp]),
        "sb2_program_text": (ci, [vp, C.c_char_p, ci]),
        "sb2_specialized": (ci, [vp]),
        "sb2_specialize_info": (ci, [vp, C.c_char_p, ci]),
        "sb2_specialize_now": (ci, [vp]),
        "sb2_specialize_dry_run": (ci, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), ci, ci, ci, ci, ci, ci, ci, cs, C.c_char_p, ci]),
        # --- include/spatial_sim_c.h
        "ssb_new": (vp, []),
        "ssb_free": (None, [vp]),
        "ssb_last_error": (cs, [vp]),
        "ssb_is_initialized": (ci, [vp]),
        "ssb_set_device": (None, [vp, ci]),
        "ssb_set_host_only": (None, [vp, ci]),
        "ssb_set_slab": (None, [vp, ci, ci]),
        "ssb_set_devices": (None, [vp, C.POINTER(ci), ci]),
        "ssb_get_num_slabs": (ci, [vp]),
        "ssb_get_slab_rows": (None, [vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]),
        "ssb_set_cuda_stream": (None, [vp, vp]),
        "ssb_init_from_file": (ci, [vp, cs, ci, ci, ci]),
        "ssb_init_from_string": (ci, [vp, cs, ci, ci, ci]),
        "ssb_one_step": (cd, [vp, cd, cd]),
        "ssb_run_steps": (ci, [vp, cd, cd, ci]),
        "ssb_one_step_host": (cd, [vp, cd, cd, ci, C.POINTER(C.c_char_p), C.POINTER(dp), C.POINTER(dp), ci]),
        "ssb_run": (ci, [vp, cd, cd]),
        "ssb_set_parameter": (None, [vp, cs, cd]),
        "ssb_set_parameter_uniformly": (None, [vp, cs, cd]),
        "ssb_get_index_for_position": (ci, [vp, cd, cd]),
        "ssb_get_variable_length": (ci, [vp]),
        "ssb_get_geometry_length": (ci, [vp]),
        "ssb_get_xdim": (ci, [vp]),
        "ssb_get_ydim": (ci, [vp]),
        "ssb_get_zdim": (ci, [vp]),
        "ssb_get_variable_at": (cd, [vp, cs, ci, ci, ci]),
        "ssb_flip_volume_order": (None, [vp]),
        "ssb_get_variable": (dp, [vp, cs, C.POINTER(ci)]),
        "ssb_release_variable": (None, [vp, cs]),
        "ssb_get_geometry": (C.POINTER(ci), [vp, cs, C.POINTER(ci)]),
        "ssb_get_boundary_type": (C.POINTER(C.c_uint8), [vp, cs, C.POINTER(ci)]),
        "ssb_get_boundary": (C.POINTER(ci), [vp, cs, C.POINTER(ci)]),
        "ssb_get_x": (dp, [vp, C.POINTER(ci)]),
        "ssb_get_y": (dp, [vp, C.POINTER(ci)]),
        "ssb_get_z": (dp, [vp, C.POINTER(ci)]),
        "ssb_get_variable_compact": (ci, [vp, cs, dp, C.c_size_t]),
        "ssb_set_variable_compact": (ci, [vp, cs, dp, C.c_size_t]),
        "ssb_get_stable_time_step": (cd, [vp, cd]),
        "ssb_get_num_domain_cells": (C.c_long, [vp]),
        "ssb_get_num_staged_species": (ci, [vp]),
        "ssb_get_time_slot": (ci, [vp]),
        "ssb_get_device_launch_count": (C.c_long, [vp]),
        "ssb_get_device_handle": (vp, [vp]),
        "ssb_device_state_changed": (None, [vp]),
        "ssb_get_device_program_text": (ci, [vp, C.c_char_p, ci]),
        "ssb_is_specialized": (ci, [vp]),
        "ssb_get_specialize_info": (ci, [vp, C.c_char_p, ci]),
        "ssb_specialize_now": (ci, [vp]),
        "ssb_get_setup_text": (ci, [vp, cs, C.c_char_p, ci]),
    }
    for name, (res, args) in P.items():
        f = getattr(lib, name)  # AttributeError if the library does not export what the headers declare
        f.restype = res
        f.argtypes = args
    return P


def load_library():
    """Load libspatialb200.so (built in-tree by build.sh / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run ./build.sh (there is no Python / CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib._prototypes = _proto(lib)
        _lib = lib
    return _lib


# opcodes of the device bytecode (enum Sb2Opc, csrc/device/sb2_internal.h; the tests compare the two)
RD_OPCODES = ["END", "PUSHC", "PUSHV", "PUSHF", "ADD", "SUB", "MUL", "DIV", "GEQ", "GT", "LEQ", "LT", "AND", "OR",
              "ADDC", "SUBC", "MULC", "DIVC", "RSUBC", "RDIVC", "GEQC", "GTC", "LEQC", "LTC", "ADDV", "SUBV", "MULV", "DIVV",
              "UNARY", "BINR", "MOV0", "MOVUP", "MASK", "ACCSUB", "ACCADD", "ACCSUB1", "ACCADD1", "ACCSUBC1", "ACCADDC1",
              "ACCSUBM", "ACCADDM", "ADDVC", "SUBVC", "RSUBVC", "MULVC", "DIVVC", "RDIVVC", "ADDVV", "SUBVV", "MULVV", "DIVVV", "XFER"]
SB2_MAX_DEPTH = 8


def specialize_dry_run(records, ns, np_, w, depth, minb=2, d3=False, cubin_path=None):
    """Builds the model-specialised stage kernel for a list of (opcode name, slot, arg) records with the CUDA
    run-time compiler, without a GPU (sb2_specialize_dry_run).  Returns (rc, compiler log)."""
    lib = load_library()
    n = len(records)
    code = (C.c_uint32 * max(n, 1))(*[RD_OPCODES.index(o) * SB2_MAX_DEPTH + d for o, d, a in records])
    arg = (C.c_uint32 * max(n, 1))(*[a for o, d, a in records])
    log = C.create_string_buffer(1 << 16)
    rc = lib.sb2_specialize_dry_run(code, arg, n, ns, np_, w, depth, minb, 1 if d3 else 0,
                                    cubin_path.encode() if cubin_path else None, log, len(log))
    return rc, log.value.decode(errors="replace")


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class SpatialSimulator(object):
    """Python face of the SpatialSimulator class (same method names as the reference's
    SpatialSBML/spatialsimulator.h:49-294, snake_case aliases not provided on purpose)."""

    def __init__(self, device=0, host_only=False, stream=None, slab=None, devices=None):
        """devices: list of CUDA ordinals (one may repeat): the grid is cut into one slab per entry and stepped by all of them
        (SpatialSimulator::setDevices); results are bit-identical to one device."""
        self._lib = load_library()
        self._h = self._lib.ssb_new()
        self._lib.ssb_set_device(self._h, int(device))
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            self._lib.ssb_set_devices(self._h, arr, len(devices))
        if host_only:
            self._lib.ssb_set_host_only(self._h, 1)
        if slab is not None:
            self._lib.ssb_set_slab(self._h, int(slab[0]), int(slab[1]))
        self._stream = stream

    def close(self):
        if self._h:
            self._lib.ssb_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self):
        return self._lib.ssb_last_error(self._h).decode()

    def _after_init(self, rc):
        if rc != 0:
            raise RuntimeError("SpatialSimulator init failed: " + self._err())
        if self._stream is not None:
            self._lib.ssb_set_cuda_stream(self._h, C.c_void_p(self._stream))

    def initFromFile(self, path, xdim, ydim, zdim=1):
        self._after_init(self._lib.ssb_init_from_file(self._h, os.fsencode(path), xdim, ydim, zdim))

    def initFromString(self, sbml, xdim, ydim, zdim=1):
        self._after_init(self._lib.ssb_init_from_string(self._h, sbml.encode(), xdim, ydim, zdim))

    def oneStep(self, initial_time, step):
        r = self._lib.ssb_one_step(self._h, float(initial_time), float(step))
        if r != r:
            raise RuntimeError("oneStep failed: " + self._err())
        return r

    def oneStepHost(self, initial_time, step, inputs=None, outputs=None, nslabs=0):
        """oneStep with species values taken from / returned to host arrays: inputs / outputs map species id → contiguous
        float64 array of localShape() (pinned memory makes the copies asynchronous).  Upload, the four RK stages and
        download are pipelined over slabs of the slowest axis."""
        inputs, outputs = inputs or {}, outputs or {}
        ids = sorted(set(inputs) | set(outputs))
        n = len(ids)
        c_ids = (C.c_char_p * n)(*[i.encode() for i in ids])
        dp = C.POINTER(C.c_double)
        c_in = (dp * n)(*[_dptr(inputs[i]) if i in inputs else dp() for i in ids])
        c_out = (dp * n)(*[_dptr(outputs[i]) if i in outputs else dp() for i in ids])
        for a in list(inputs.values()) + list(outputs.values()):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == int(np.prod(self.localShape()))
        r = self._lib.ssb_one_step_host(self._h, float(initial_time), float(step), n, c_ids, c_in, c_out, int(nslabs))
        if r != r:
            raise RuntimeError("oneStepHost failed: " + self._err())
        return r

    def runSteps(self, first_step_index, step, nsteps):
        if self._lib.ssb_run_steps(self._h, float(first_step_index), float(step), int(nsteps)) != 0:
            raise RuntimeError("runSteps failed: " + self._err())

    def run(self, end_time, step):
        if self._lib.ssb_run(self._h, float(end_time), float(step)) != 0:
            raise RuntimeError("run failed: " + self._err())

    def setParameter(self, id, value):
        self._lib.ssb_set_parameter(self._h, id.encode(), float(value))

    def setParameterUniformly(self, id, value):
        self._lib.ssb_set_parameter_uniformly(self._h, id.encode(), float(value))

    def getIndexForPosition(self, x, y):
        return self._lib.ssb_get_index_for_position(self._h, float(x), float(y))

    def getVariableLength(self):
        return self._lib.ssb_get_variable_length(self._h)

    def getGeometryLength(self):
        return self._lib.ssb_get_geometry_length(self._h)

    def getXDim(self):
        return self._lib.ssb_get_xdim(self._h)

    def getYDim(self):
        return self._lib.ssb_get_ydim(self._h)

    def getZDim(self):
        return self._lib.ssb_get_zdim(self._h)

    def getVariableAt(self, id, x, y, z=0):
        return self._lib.ssb_get_variable_at(self._h, id.encode(), int(x), int(y), int(z))

    def flipVolumeOrder(self):
        self._lib.ssb_flip_volume_order(self._h)

    def _borrow(self, fn, key, dtype, width=1):
        n = C.c_int(0)
        p = fn(self._h, key.encode(), C.byref(n)) if key is not None else fn(self._h, C.byref(n))
        if not p:
            return None, n.value
        return p, n.value

    def getVariable(self, id):
        """Live doubled-grid view (numpy array over the library's buffer; length+1 entries are backed)."""
        p, n = self._borrow(self._lib.ssb_get_variable, id, np.float64)
        if p is None:
            return None
        total = self.getGeometryLength() if n > 0 and n + 1 == self.getGeometryLength() else n
        return np.ctypeslib.as_array(p, shape=(total,))

    def getSpeciesCompartment(self, id):
        """compartment id of a variable (what getModel()->getSpecies(id)->getCompartment() gives a caller of the reference), or None"""
        for line in self.getSetupText("summary").split("\n"):
            t = line.split()
            if len(t) > 2 and t[0] == "var" and t[1] == id and "geo" in t:
                c = t[t.index("geo") + 1]
                return None if c == "-" else c
        return None

    def releaseVariable(self, id):
        """Stop keeping the array behind getVariable(id) current after every step (a download per step otherwise)."""
        self._lib.ssb_release_variable(self._h, id.encode())

    def getGeometry(self, compartment):
        p, n = self._borrow(self._lib.ssb_get_geometry, compartment, np.int32)
        return None if p is None else np.ctypeslib.as_array(p, shape=(n,))

    def getBoundary(self, compartment):
        p, n = self._borrow(self._lib.ssb_get_boundary, compartment, np.int32)
        return None if p is None else np.ctypeslib.as_array(p, shape=(n,))

    def getBoundaryType(self, compartment):
        p, n = self._borrow(self._lib.ssb_get_boundary_type, compartment, np.uint8)
        return None if p is None else np.ctypeslib.as_array(p, shape=(n, 6))

    def getX(self):
        p, n = self._borrow(self._lib.ssb_get_x, None, np.float64)
        return None if p is None else np.ctypeslib.as_array(p, shape=(n + 1,))

    def getY(self):
        p, n = self._borrow(self._lib.ssb_get_y, None, np.float64)
        return None if p is None else np.ctypeslib.as_array(p, shape=(n + 1,))

    def getZ(self):
        p, n = self._borrow(self._lib.ssb_get_z, None, np.float64)
        return None if p is None else np.ctypeslib.as_array(p, shape=(n + 1,))

    # ---- additions ----
    def localShape(self):
        """(nz, ny, nx) of the compact arrays held by this simulator (the slab incl. halo when sharded)."""
        txt = self.getSetupText("summary").split("\n")[0].split()
        return int(txt[3]), int(txt[2]), int(txt[1])

    def getVariableCompact(self, id, out=None):
        nz, ny, nx = self.localShape()
        if out is None:
            out = np.empty((nz, ny, nx), dtype=np.float64)
        if self._lib.ssb_get_variable_compact(self._h, id.encode(), _dptr(out), out.size) != 0:
            raise KeyError("no spatial variable %r" % id)
        return out

    def setVariableCompact(self, id, values):
        a = np.ascontiguousarray(values, dtype=np.float64)
        if self._lib.ssb_set_variable_compact(self._h, id.encode(), _dptr(a), a.size) != 0:
            raise KeyError("cannot set variable %r" % id)

    def getStableTimeStep(self, dt):
        r = self._lib.ssb_get_stable_time_step(self._h, float(dt))
        if r != r:
            raise RuntimeError("getStableTimeStep failed: " + self._err())
        return r

    def getNumDomainCells(self):
        return self._lib.ssb_get_num_domain_cells(self._h)

    def getNumStagedSpecies(self):
        return self._lib.ssb_get_num_staged_species(self._h)

    def getSlabRows(self):
        """(first row held, first owned row, one past the last owned row) in rows of the whole grid; zeros unless a slab"""
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self._lib.ssb_get_slab_rows(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def getNumSlabs(self):
        return self._lib.ssb_get_num_slabs(self._h)

    def getTimeSlot(self):
        """Constants-table slot of the model's `time` symbol (what sb2_begin_step takes), -1 when unused."""
        return self._lib.ssb_get_time_slot(self._h)

    def getDeviceLaunchCount(self):
        return self._lib.ssb_get_device_launch_count(self._h)

    def getDeviceHandle(self):
        return self._lib.ssb_get_device_handle(self._h)

    def deviceStateChanged(self):
        self._lib.ssb_device_state_changed(self._h)

    def setCudaStream(self, stream):
        self._stream = stream
        self._lib.ssb_set_cuda_stream(self._h, C.c_void_p(stream))

    def _text(self, fn, *args):
        n = fn(self._h, *args, None, 0)
        buf = C.create_string_buffer(n + 1)
        fn(self._h, *args, buf, n + 1)
        return buf.value.decode()

    def isSpecialized(self):
        """True when the stage kernel was built for this model's kinetics at set-up (NVRTC); the results are the same either way."""
        return bool(self._lib.ssb_is_specialized(self._h))

    def specializeNow(self):
        """Builds (or waits for the background build of) the model-specialised stage kernel; True when it is in place."""
        return self._lib.ssb_specialize_now(self._h) == 1

    def getSpecializeInfo(self):
        return self._text(self._lib.ssb_get_specialize_info)

    def getDeviceProgramText(self):
        return self._text(self._lib.ssb_get_device_program_text)

    def getSetupText(self, what):
        return self._text(self._lib.ssb_get_setup_text, what.encode())
