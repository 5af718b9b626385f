"""Host-side mirror of the reference's solver plugin interface on top of the
C ABI of ``libgfmd_b200.so`` (``include/gfmd_b200.h``).

``GFMDSolverB200`` follows ``class GFMDSolver`` (reference
``src/main/gfmd_solver.h:34-123``) member for member -- ``set_grid_size``,
``set_kernel``, ``pre_force``, ``post_force``, ``get_u0``, ``get_xlo_loc`` ... --
so the parity tests read like drivers of the reference's own
``GFMDSolverStatic``.  Inside LAMMPS the same ABI is bound from C++
(``user-gfmd_b200/host/gfmd_solver_b200.{h,cpp}``).

There is no CPU fallback: if the CUDA library is missing or no GPU is usable,
loading / creating raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libgfmd_b200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
NSTAGES = 9
STAGE_NAMES = ("gather", "rows_fwd", "exchange_fwd", "cols_fused", "exchange_inv", "rows_inv",
               "scatter", "cols_top_fwd", "cols_top_inv")
UNIQUE_ID_BYTES = 128
IPC_HANDLE_BYTES = 64

# every symbol include/gfmd_b200.h declares: (restype, argtypes)
_vp = ctypes.c_void_p
_i = ctypes.c_int
_d = ctypes.c_double
ABI = {
    "gfmd_b200_create": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i]),
    "gfmd_b200_create_slab": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i]),
    "gfmd_b200_get_unique_id": (_i, [ctypes.c_char_p]),
    "gfmd_b200_comm_init": (_i, [_vp, ctypes.c_char_p]),
    "gfmd_b200_ipc_export": (_i, [_vp, ctypes.c_char_p]),
    "gfmd_b200_ipc_import": (_i, [_vp, ctypes.c_char_p]),
    "gfmd_b200_ipc_export_stage": (_i, [_vp, ctypes.c_char_p]),
    "gfmd_b200_ipc_import_stage": (_i, [_vp, ctypes.c_char_p]),
    "gfmd_b200_destroy": (None, [_vp]),
    "gfmd_b200_last_error": (ctypes.c_char_p, [_vp]),
    "gfmd_b200_get_brick": (_i, [_vp, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "gfmd_b200_get_q_columns": (_i, [_vp, c_int_p, c_int_p]),
    "gfmd_b200_set_phi": (_i, [_vp, _vp, _i, _vp]),
    "gfmd_b200_set_phi_columns": (_i, [_vp, _vp, _i, _i, _i]),
    "gfmd_b200_set_linf": (_i, [_vp, _vp]),
    "gfmd_b200_build_phi_columns": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "gfmd_b200_build_phi_columns_device": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "gfmd_b200_phi_deviation": (_i, [_vp, c_double_p, c_double_p]),
    "gfmd_b200_post_force_host": (_i, [_vp, _vp, _vp, c_double_p, _vp]),
    "gfmd_b200_pre_force_async_host": (_i, [_vp, _vp]),
    "gfmd_b200_post_force_device": (_i, [_vp, _vp, _vp]),
    "gfmd_b200_spectrum_host": (_i, [_vp, _vp, _vp, _vp]),
    "gfmd_b200_prec_gradient_host": (_i, [_vp, _vp, _vp, _vp, _i]),
    "gfmd_b200_gather": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _d, _d, _i, _i, _vp]),
    "gfmd_b200_scatter": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "gfmd_b200_build_cell_map": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, c_int_p]),
    "gfmd_b200_drop_cell_map": (_i, [_vp]),
    "gfmd_b200_full_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _d, _d, _vp]),
    "gfmd_b200_get_results": (_i, [_vp, c_double_p, _vp, _vp, _vp]),
    "gfmd_b200_device_u": (_vp, [_vp]),
    "gfmd_b200_device_f": (_vp, [_vp]),
    "gfmd_b200_stream": (_vp, [_vp]),
    "gfmd_b200_set_stream": (_i, [_vp, _vp]),
    "gfmd_b200_synchronize": (_i, [_vp]),
    "gfmd_b200_pin_host_buffers": (_i, [_vp, _i]),
    "gfmd_b200_use_graph": (_i, [_vp, _i]),
    "gfmd_b200_host_pipeline": (_i, [_vp, _i]),
    "gfmd_b200_launch_count": (ctypes.c_longlong, [_vp]),
    "gfmd_b200_profile": (_i, [_vp, _i]),
    "gfmd_b200_get_stage_times": (_i, [_vp, _vp, _vp]),
    "gfmd_b200_describe": (ctypes.c_char_p, [_vp]),
    "gfmd_b200_memory_usage": (_d, [_vp]),
    "gfmd_b200_version": (ctypes.c_char_p, []),
}

_lib = None


class GFMDError(RuntimeError):
    """Mirrors LAMMPS error->one(FLERR, msg): the reference's only error path."""

    def __init__(self, code, msg):
        super().__init__("gfmd_b200 error %d: %s" % (code, msg))
        self.code = code


def load_library(path=None):
    """Loads libgfmd_b200.so; raises (no fallback) if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # GFMD_B200_LIB: another build of the SAME library (A/B of compile-time options, tools/ only)
    p = path or os.environ.get("GFMD_B200_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(user-gfmd_b200/csrc/build.sh). There is no CPU fallback." % p)
    lib = ctypes.CDLL(p)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)      # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _ptr(a):
    """Raw address of a numpy array, torch tensor (host or device) or int."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError("cannot take the address of %r" % type(a))


def get_unique_id():
    buf = ctypes.create_string_buffer(UNIQUE_ID_BYTES)
    rc = load_library().gfmd_b200_get_unique_id(buf)
    if rc:
        raise GFMDError(rc, load_library().gfmd_b200_last_error(None).decode())
    return buf.raw


class GFMDSolverB200:
    """``static/b200``: drop-in for ``GFMDSolverStatic`` behind ``GFMDSolver``.

    Reference interface (gfmd_solver.h) -> here:
      GFMDSolver(LAMMPS*)                 -> GFMDSolverB200(device=0, rank=0, nranks=1)
      set_grid_size(nx, ny, ndof)  (:41)  -> set_grid_size
      set_kernel(kernel, normalize) (:42) -> set_kernel(phi_table, linf, normalized)
                                             (the table fill_phi_buffer produced)
      pre_force(u, f)              (:50)  -> pre_force   (asynchronous launch)
      post_force(u, f, dump_prefix) (:51) -> post_force  -> epot of this rank
      get_u0 / get_xlo_loc ...     (:72-90)
      get_name                     (:92)  -> "static/b200"
      memory_usage                 (:96)
    """

    name = "static/b200"

    def __init__(self, device=0, rank=0, nranks=1, unique_id=None):
        self.lib = load_library()
        self.h = None
        self.device, self.rank, self.nranks = device, rank, nranks
        self._unique_id = unique_id
        self.u0 = None
        self.peer_stage_error = None      # why the optional row-buffer mapping is missing, if it is

    # -- reference interface ------------------------------------------------
    def get_name(self):
        return self.name

    def set_grid_size(self, nx, ny, ndof):
        if self.h:
            self.close()
        h = ctypes.c_void_p()
        if self.nranks == 1:
            rc = self.lib.gfmd_b200_create(ctypes.byref(h), nx, ny, ndof, self.device)
        else:
            rc = self.lib.gfmd_b200_create_slab(ctypes.byref(h), nx, ny, ndof, self.device,
                                                self.rank, self.nranks)
        if rc:
            raise GFMDError(rc, self.lib.gfmd_b200_last_error(None).decode())
        self.h = h
        self.nx, self.ny, self.ndof = nx, ny, ndof
        self.u0 = np.zeros(ndof)
        # NCCL is optional: with peer mappings (enable_peer_copy) the ranks order their transfers
        # through flag words in each other's memory and no communicator is needed
        if self.nranks > 1 and self._unique_id is not None:
            self._check(self.lib.gfmd_b200_comm_init(self.h, self._unique_id))
        v = [ctypes.c_int() for _ in range(6)]
        self._check(self.lib.gfmd_b200_get_brick(self.h, *[ctypes.byref(x) for x in v]))
        (self.xlo_loc, self.xhi_loc, self.ylo_loc, self.yhi_loc, self.nxy_loc,
         self.gammai) = [x.value for x in v]
        a, b = ctypes.c_int(), ctypes.c_int()
        self._check(self.lib.gfmd_b200_get_q_columns(self.h, ctypes.byref(a), ctypes.byref(b)))
        self.kylo, self.nky = a.value, b.value

    def init(self):
        pass

    def ipc_export(self):
        """This rank's two receive-buffer handles (128 bytes) for the peer-copy exchange."""
        buf = ctypes.create_string_buffer(2 * IPC_HANDLE_BYTES)
        self._check(self.lib.gfmd_b200_ipc_export(self.h, buf))
        return buf.raw

    def ipc_import(self, all_handles):
        """all_handles: the exports of all ranks concatenated in rank order."""
        assert len(all_handles) == 2 * IPC_HANDLE_BYTES * self.nranks
        self._check(self.lib.gfmd_b200_ipc_import(self.h, all_handles))

    def ipc_export_stage(self):
        """Handle (64 bytes) of the buffer this rank's row kernels write (GFMD_B200_PEER_DIRECT)."""
        buf = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(self.lib.gfmd_b200_ipc_export_stage(self.h, buf))
        return buf.raw

    def ipc_import_stage(self, all_handles):
        assert len(all_handles) == IPC_HANDLE_BYTES * self.nranks
        self._check(self.lib.gfmd_b200_ipc_import_stage(self.h, all_handles))

    def enable_peer_copy(self, all_gather_bytes):
        """Convenience: all_gather_bytes(b) -> list of every rank's bytes, in rank order.  Maps the
        peers' receive buffers and their row-output buffers (the latter is only used with
        GFMD_B200_PEER_DIRECT=1)."""
        self.ipc_import(b"".join(all_gather_bytes(self.ipc_export())))
        # every rank takes part in the gather; a failed optional mapping only disables the direct mode
        try:
            mine = self.ipc_export_stage()
        except GFMDError as e:
            mine, self.peer_stage_error = b"\0" * IPC_HANDLE_BYTES, str(e)
        everyone = all_gather_bytes(mine)
        if any(h == b"\0" * IPC_HANDLE_BYTES for h in everyone):
            return
        try:
            self.ipc_import_stage(b"".join(everyone))
        except GFMDError as e:
            self.peer_stage_error = str(e)

    def set_kernel(self, phi, linf=None, normalized=True):
        """phi: the table fill_phi_buffer produced for the whole grid,
        [nx*ny, ndof, ndof] complex128 (gfmd_misc.cpp:32-101)."""
        phi = np.ascontiguousarray(phi, dtype=np.complex128)
        if phi.size != self.nx * self.ny * self.ndof * self.ndof:
            raise GFMDError(1, "phi table has %d entries, expected %d" %
                            (phi.size, self.nx * self.ny * self.ndof ** 2))
        lp = None
        if linf is not None:
            linf = np.ascontiguousarray(linf, dtype=np.float64)
            lp = linf.ctypes.data
        self._check(self.lib.gfmd_b200_set_phi(self.h, phi.ctypes.data, int(normalized), lp))

    def set_kernel_columns(self, phi_cols, ky_first, normalized=True):
        """phi_cols: [nx, nky, ndof, ndof] as fill_phi_buffer(ndof, nx, 0, nx-1, ny,
        ky_first, ky_first+nky-1) lays it out."""
        phi_cols = np.ascontiguousarray(phi_cols, dtype=np.complex128)
        nky = phi_cols.size // (self.nx * self.ndof * self.ndof)
        self._check(self.lib.gfmd_b200_set_phi_columns(self.h, phi_cols.ctypes.data, ky_first, nky,
                                                       int(normalized)))

    def build_kernel_columns(self, uuv, ky_first, height, normalize=True):
        """Device-side transfer-matrix recursion.  uuv: [nx, nky, 3, ndof, ndof] complex128 =
        (U0, U, V) of StiffnessKernel::get_dynamical_matrices for every q of the column block."""
        uuv = np.ascontiguousarray(uuv, dtype=np.complex128)
        nky = uuv.size // (self.nx * 3 * self.ndof * self.ndof)
        self._check(self.lib.gfmd_b200_build_phi_columns(self.h, uuv.ctypes.data, ky_first, nky, height,
                                                         int(normalize)))

    def build_kernel_columns_device(self, d_uuv, ky_first, nky, height, normalize=True):
        """build_kernel_columns with the (U0, U, V) blocks in device memory (torch tensor or address)."""
        self._check(self.lib.gfmd_b200_build_phi_columns_device(self.h, _ptr(d_uuv), ky_first, nky, height,
                                                                int(normalize)))

    def set_linf(self, linf):
        linf = np.ascontiguousarray(linf, dtype=np.float64)
        self._check(self.lib.gfmd_b200_set_linf(self.h, linf.ctypes.data))

    def pre_force(self, u, f=None):
        """Asynchronous start (gfmd_solver.h:44-50).  u: host [ndof, nxy_loc]."""
        self._check(self.lib.gfmd_b200_pre_force_async_host(self.h, _ptr(u)))

    def post_force(self, u, f, dump_prefix=None):
        """u, f: HOST arrays [ndof, nxy_loc] (numpy or pinned torch).  Fills f and
        u0, returns this rank's potential energy (gfmd_solver_static.cpp:145-249)."""
        if dump_prefix is not None:
            raise GFMDError(4, "q-space dumps are written by the host shim, not by this binding")
        for a in (u, f):
            if isinstance(a, np.ndarray) and (a.dtype != np.float64 or not a.flags.c_contiguous):
                raise GFMDError(1, "u and f must be C-contiguous float64")
        e = ctypes.c_double()
        self._check(self.lib.gfmd_b200_post_force_host(self.h, _ptr(u), _ptr(f), ctypes.byref(e),
                                                       self.u0.ctypes.data))
        return e.value

    def prec_gradient(self, cavg, g, gP, reference_quirk=True):
        """GFMDSolverStatic::prec_gradient (gfmd_solver_static.cpp:253-271):
        gP = IDFT[(Phi(q) + cavg)^-1 DFT[g]].  cavg: [ndof, ndof] real; g, gP: HOST [ndof, nxy_loc].
        reference_quirk: for ndof > 3 only components 0..2 of gP(q) are replaced, like the
        reference's general branch (gfmd_misc.h:113-115); False replaces all."""
        cavg = np.ascontiguousarray(cavg, dtype=np.float64)
        if cavg.size != self.ndof * self.ndof:
            raise GFMDError(1, "cavg must hold ndof*ndof entries")
        for a in (g, gP):
            if isinstance(a, np.ndarray) and (a.dtype != np.float64 or not a.flags.c_contiguous):
                raise GFMDError(1, "g and gP must be C-contiguous float64")
        self._check(self.lib.gfmd_b200_prec_gradient_host(self.h, cavg.ctypes.data, _ptr(g), _ptr(gP),
                                                          int(reference_quirk)))

    def spectrum(self, u, with_force=True):
        """The two q-space fields GFMDSolverFFT::dump writes (gfmd_solver_fft.cpp:209-287):
        u~(q) and Phi(q).u~(q), each [nx*ny, ndof] complex128 (idq = ix*ny + iy)."""
        if isinstance(u, np.ndarray) and (u.dtype != np.float64 or not u.flags.c_contiguous):
            raise GFMDError(1, "u must be C-contiguous float64")
        uq = np.empty((self.nx * self.ny, self.ndof), dtype=np.complex128)
        fq = np.empty_like(uq) if with_force else None
        self._check(self.lib.gfmd_b200_spectrum_host(self.h, _ptr(u), uq.ctypes.data,
                                                     fq.ctypes.data if with_force else None))
        return uq, fq

    def get_u0(self):
        return self.u0

    def get_xlo_loc(self):
        return self.xlo_loc

    def get_xhi_loc(self):
        return self.xhi_loc

    def get_ylo_loc(self):
        return self.ylo_loc

    def get_yhi_loc(self):
        return self.yhi_loc

    def get_nxy_loc(self):
        return self.nxy_loc

    def memory_usage(self):
        return self.lib.gfmd_b200_memory_usage(self.h)

    # -- device-resident path -------------------------------------------------
    def post_force_device(self, d_u=None, d_f=None):
        self._check(self.lib.gfmd_b200_post_force_device(self.h, _ptr(d_u), _ptr(d_f)))

    def gather(self, d_x, d_xeq, d_gid, d_mask, groupbit, nall, xprd, yprd, dxshift=0, dyshift=0,
               d_u=None):
        self._check(self.lib.gfmd_b200_gather(self.h, _ptr(d_x), _ptr(d_xeq), _ptr(d_gid),
                                              _ptr(d_mask), groupbit, nall, xprd, yprd, dxshift,
                                              dyshift, _ptr(d_u)))

    def scatter(self, d_gid, d_mask, groupbit, nall, nlocal, d_f, d_fgrid=None):
        self._check(self.lib.gfmd_b200_scatter(self.h, _ptr(d_fgrid), _ptr(d_gid), _ptr(d_mask),
                                               groupbit, nall, nlocal, _ptr(d_f)))

    def build_cell_map(self, d_gid, d_mask, groupbit, nall, nlocal, dxshift=0, dyshift=0):
        """Cell -> atom map for the fused gather / scatter form of full_step; returns whether it is
        usable (one atom per cell and radix-16 row kernels), see include/gfmd_b200.h."""
        ok = ctypes.c_int()
        self._check(self.lib.gfmd_b200_build_cell_map(self.h, _ptr(d_gid), _ptr(d_mask), groupbit, nall, nlocal,
                                                      dxshift, dyshift, ctypes.byref(ok)))
        return bool(ok.value)

    def drop_cell_map(self):
        self._check(self.lib.gfmd_b200_drop_cell_map(self.h))

    def full_step(self, d_x, d_xeq, d_gid, d_mask, groupbit, nall, nlocal, xprd, yprd, d_f):
        self._check(self.lib.gfmd_b200_full_step(self.h, _ptr(d_x), _ptr(d_xeq), _ptr(d_gid),
                                                 _ptr(d_mask), groupbit, nall, nlocal, xprd, yprd,
                                                 _ptr(d_f)))

    def results(self):
        """Synchronises; returns dict(epot, u0, fsum, natoms_gathered, natoms_scattered,
        n_out_of_range)."""
        e = ctypes.c_double()
        fsum = np.zeros(3)
        cnt = np.zeros(3, dtype=np.int32)
        self._check(self.lib.gfmd_b200_get_results(self.h, ctypes.byref(e), self.u0.ctypes.data,
                                                   fsum.ctypes.data, cnt.ctypes.data))
        return dict(epot=e.value, u0=self.u0.copy(), fsum=fsum, natoms_gathered=int(cnt[0]),
                    natoms_scattered=int(cnt[1]), n_out_of_range=int(cnt[2]))

    # -- plumbing ---------------------------------------------------------------
    def device_u(self):
        return self.lib.gfmd_b200_device_u(self.h)

    def device_f(self):
        return self.lib.gfmd_b200_device_f(self.h)

    def stream(self):
        return self.lib.gfmd_b200_stream(self.h)

    def set_stream(self, cuda_stream):
        self._check(self.lib.gfmd_b200_set_stream(self.h, cuda_stream))

    def synchronize(self):
        self._check(self.lib.gfmd_b200_synchronize(self.h))

    def pin_host_buffers(self, on=True):
        self._check(self.lib.gfmd_b200_pin_host_buffers(self.h, int(on)))

    def host_pipeline(self, on=None):
        """Per-dof upload / download pipeline of the host path (see include/gfmd_b200.h);
        on=None only queries.  Returns the setting."""
        rc = self.lib.gfmd_b200_host_pipeline(self.h, -1 if on is None else int(bool(on)))
        if rc > 1:
            self._check(rc)
        return bool(rc)

    def use_graph(self, on=True):
        self._check(self.lib.gfmd_b200_use_graph(self.h, int(on)))

    def launch_count(self):
        return self.lib.gfmd_b200_launch_count(self.h)

    def profile(self, on=True):
        self._check(self.lib.gfmd_b200_profile(self.h, int(on)))

    def stage_times(self):
        ms = np.zeros(NSTAGES)
        cnt = np.zeros(NSTAGES, dtype=np.int64)
        self._check(self.lib.gfmd_b200_get_stage_times(self.h, ms.ctypes.data, cnt.ctypes.data))
        return {n: (float(ms[i]), int(cnt[i])) for i, n in enumerate(STAGE_NAMES)}

    def describe(self):
        return self.lib.gfmd_b200_describe(self.h).decode()

    def phi_deviation(self):
        a, b = ctypes.c_double(), ctypes.c_double()
        self._check(self.lib.gfmd_b200_phi_deviation(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def close(self):
        if self.h:
            self.lib.gfmd_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise GFMDError(rc, self.lib.gfmd_b200_last_error(self.h).decode())


def gfmd_solver_factory(keyword=None, **kw):
    """Mirror of gfmd_solver_factory (src/main/gfmd_solver.cpp:209-257): only the
    B200 solver exists here; the name check of :244-253 is kept."""
    name = keyword or "static/b200"
    if name not in ("static/b200", "b200"):
        raise GFMDError(1, "Unknown solver name encountered.")
    return GFMDSolverB200(**kw)


def all_gather_bytes_fn(dev, world):
    import torch
    import torch.distributed as dist

    def f(b):
        # dev: where the process group's backend wants its tensors (cuda for nccl, cpu for gloo)
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [bytes(o.cpu().numpy().tobytes()) for o in out]
    return f


def slab_plan(nx, ny, ndof, rank, nranks):
    """The x-slab / ky-slab decomposition the library uses (create_common in
    csrc/gfmd_b200.cu), for hosts that place atoms and stiffness tables per rank.

    Real space: rank r owns rows [x0, x0 + nx_loc) (the reference's brick for
    procgrid = P x 1 x 1, src/main/gfmd_solver.cpp:95-101).  After the transpose it owns
    the half-spectrum columns ky in [ky0, ky0 + nky_loc), all kx.  The exchange buffers are
    [P][ndof][kyb][nx_loc] complex: block p is what goes to / comes from rank p."""
    if nx % nranks:
        raise GFMDError(4, "nx = %d not divisible by %d slab ranks" % (nx, nranks))
    nyh = ny // 2 + 1
    kyb = (nyh + nranks - 1) // nranks
    ky0 = rank * kyb
    nky_loc = max(0, min(kyb, nyh - ky0))
    nx_loc = nx // nranks
    return dict(nx_loc=nx_loc, x0=rank * nx_loc, nyh=nyh, kyb=kyb, ky0=ky0, nky_loc=nky_loc,
                block_elems=ndof * kyb * nx_loc, gamma_rank=0,
                ky_weights=[1.0 if (k == 0 or 2 * k == ny) else 2.0 for k in range(ky0, ky0 + nky_loc)])
