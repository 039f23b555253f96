"""Thin object wrapper over a ``dlra_handle`` (include/dlra.h).  Holds no numerics: every method is one C-ABI call.
Device memory handed to the engine is either a NumPy array (host path, ``*_host`` entry points) or a CUDA
``torch.Tensor`` used purely as a device allocation (PyTorch is plumbing here, not the compute path)."""
import ctypes as C

import numpy as np

from . import _lib as L

try:  # torch is only needed for device-resident inputs and torch.distributed plumbing
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def colmajor_device(x):
    """Return a CUDA fp64 tensor with column-major (Julia) layout, i.e. stride (1, ld); copies only if needed."""
    assert _is_torch(x) and x.is_cuda and x.dtype == torch.float64 and x.dim() == 2
    if x.shape[0] == 1 or x.stride(0) == 1:
        if x.shape[1] == 1 or x.stride(1) >= x.shape[0]:
            return x
    return x.t().contiguous().t()


def empty_colmajor(n, m, device):
    return torch.empty((m, n), dtype=torch.float64, device=device).t()


def _ptr_ld(x):
    """(pointer, ld, is_host, keepalive) of a column-major fp64 matrix."""
    if _is_torch(x):
        if not x.is_cuda:
            x = x.numpy()
        else:
            x = colmajor_device(x)
            ld = x.stride(1) if x.shape[1] > 1 else x.shape[0]
            return x.data_ptr(), int(ld), False, x
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if not a.flags.f_contiguous:
        a = np.asfortranarray(a)
    return a.ctypes.data, int(a.shape[0]), True, a


class Engine:
    def __init__(self, n_local, m, r0, rmax=None, rank_adaptive=False, device=None, force_generic=False, aug_basis_first=False):
        self.lib = L.load()
        if torch is not None and torch.cuda.is_available():
            device = torch.cuda.current_device() if device is None else device
        elif device is None:
            device = 0
        self.device = int(device)
        self.n, self.m = int(n_local), int(m)
        rmax = r0 if rmax is None else rmax
        flags = ((L.RANK_ADAPTIVE if rank_adaptive else 0) | (L.FORCE_GENERIC if force_generic else 0)
                 | (L.AUG_BASIS_FIRST if aug_basis_first else 0))
        h = L.handle_t()
        rc = self.lib.dlra_create(self.device, self.n, self.m, int(r0), int(rmax), flags, C.byref(h))
        if rc != L.OK:
            raise L.DLRAError(rc, (self.lib.dlra_last_error(None) or b"").decode())
        self.h = h
        self.rmax = int(rmax)
        self._keep = {}
        # borrowed device snapshots: (index of the step that reads them last, tensor), released by GPU progress (dlra_progress)
        self._held = []
        self._npush = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.dlra_destroy(self.h)   # synchronises the engine's streams
            self.h = None
            self._held = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        L.check(self.h, rc)

    MAX_RUN_AHEAD = 8   # steps the host may enqueue beyond the device's progress while it feeds borrowed device snapshots

    def progress(self, wait_for=-1):
        """(steps enqueued, steps completed on the device); blocks until `wait_for` steps are complete if given."""
        enq, done = C.c_int64(), C.c_int64()
        self._ck(self.lib.dlra_progress(self.h, C.byref(enq), C.byref(done), int(wait_for)))
        return enq.value, done.value

    def _borrow(self, keep, last_reader_step):
        """A device tensor the engine reads asynchronously (dlra.h: a borrowed snapshot must stay valid until the step after the
        next push has RUN).  The steps are asynchronous and the host may be several steps ahead, so the tensor is held here
        until the device has completed its last reader; torch's caching allocator can then not hand its memory to a
        later y(t) while a queued step still reads it.  Also bounds the host's run-ahead (and with it the held memory)."""
        if not (_is_torch(keep) and keep.is_cuda):
            return
        self._held.append((last_reader_step, keep))
        enq, done = self.progress()
        if enq - done > self.MAX_RUN_AHEAD:
            enq, done = self.progress(wait_for=enq - self.MAX_RUN_AHEAD)
        while self._held and self._held[0][0] <= done:
            self._held.pop(0)

    def _after_torch(self):
        """Device tensors are produced on torch's current stream, the engine runs on its own: make the engine stream wait
        for everything torch has enqueued so far (no host blocking)."""
        if torch is not None and torch.cuda.is_available():
            self._ck(self.lib.dlra_wait_stream(self.h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    # -- multi-GPU ---------------------------------------------------------------------------------
    def comm_init(self, nranks, rank, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.dlra_comm_init(self.h, nranks, rank, buf))

    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.dlra_p2p_export(self.h, buf))
        return buf.raw

    def p2p_import(self, nranks, rank, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * nranks
        buf = C.create_string_buffer(blob, len(blob))
        self._ck(self.lib.dlra_p2p_import(self.h, nranks, rank, buf))

    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = L.load()
        buf = C.create_string_buffer(128)
        rc = lib.dlra_nccl_unique_id(buf)
        if rc != L.OK:
            raise L.DLRAError(rc, (lib.dlra_last_error(None) or b"").decode())
        return buf.raw

    # -- factors -----------------------------------------------------------------------------------
    @property
    def rank(self):
        r = C.c_int()
        self._ck(self.lib.dlra_get_rank(self.h, C.byref(r)))
        return r.value

    def set_factors(self, U, S, V):
        r = int(np.shape(S)[0]) if not _is_torch(S) else int(S.shape[0])
        pu, ldu, hu, ku = _ptr_ld(U)
        ps, lds, hs, ks = _ptr_ld(S)
        pv, ldv, hv, kv = _ptr_ld(V)
        assert hu == hs == hv, "U, S, V must all live on the host or all on the device"
        fn = self.lib.dlra_set_factors_host if hu else self.lib.dlra_set_factors
        if not hu:
            self._after_torch()
        self._ck(fn(self.h, pu, ldu, ps, lds, pv, ldv, r))
        if not hu:
            self.sync()

    def truncated_svd(self, A, r=0, tol=0.0, oversample=8, power_iters=2, seed=0):
        """Set the engine's factors to a rank-r (or tolerance-selected) truncated SVD of the device matrix A (randomized
        subspace iteration inside libdlra.so; A is streamed 2*(power_iters+1) times)."""
        p, ld, host, keep = _ptr_ld(A)
        assert not host, "truncated_svd takes a device matrix (use api.truncated_svd for host arrays)"
        self._after_torch()
        self._ck(self.lib.dlra_truncated_svd(self.h, p, ld, int(r or 0), float(tol or 0.0), int(oversample), int(power_iters), int(seed)))
        return self.rank

    def get_factors(self):
        r = self.rank
        U = np.empty((self.n, r), order="F")
        S = np.empty((r, r), order="F")
        V = np.empty((self.m, r), order="F")
        rr = C.c_int()
        self._ck(self.lib.dlra_get_factors_host(self.h, U.ctypes.data, self.n, S.ctypes.data, r, V.ctypes.data, self.m, C.byref(rr)))
        return U, S, V

    def save_factors_async(self):
        """Asynchronous snapshot of the current factors into pinned host memory (dlra_save_factors_async).  Returns
        (U, S, V) NumPy views (column-major) that are valid after save_wait(); the step stream is not stalled."""
        r = self.rank
        bufs = []
        for rows in (self.n, r, self.m):
            t = torch.empty((r, rows), dtype=torch.float64, pin_memory=True)   # row-major (r, rows) == column-major (rows, r)
            bufs.append(t)
        rr = C.c_int()
        self._ck(self.lib.dlra_save_factors_async(self.h, bufs[0].data_ptr(), self.n, bufs[1].data_ptr(), r, bufs[2].data_ptr(), self.m,
                                                  C.byref(rr)))
        return tuple(b.numpy().T for b in bufs)

    def save_wait(self):
        self._ck(self.lib.dlra_save_wait(self.h))

    def get_factors_device(self):
        r = self.rank
        dev = torch.device("cuda", self.device)
        U, S, V = empty_colmajor(self.n, r, dev), empty_colmajor(r, r, dev), empty_colmajor(self.m, r, dev)
        rr = C.c_int()
        self._ck(self.lib.dlra_get_factors(self.h, U.data_ptr(), self.n, S.data_ptr(), r, V.data_ptr(), self.m, C.byref(rr)))
        return U, S, V

    # -- data feed ---------------------------------------------------------------------------------
    def data_init(self, A0):
        p, ld, host, keep = _ptr_ld(A0)
        fn = self.lib.dlra_data_init_host if host else self.lib.dlra_data_init
        if not host:
            self._after_torch()
        self._ck(fn(self.h, p, ld))
        if host:
            self.sync_copies = True
        self._held.clear()
        self._npush = self.progress()[0]       # pushes are counted from the current step index
        self._keep["prev"] = keep
        self._borrow(keep, self._npush + 1)    # yprev is read by the next step only

    def data_push(self, A, kind=L.DATA_SNAPSHOT):
        p, ld, host, keep = _ptr_ld(A)
        fn = self.lib.dlra_data_push_host if host else self.lib.dlra_data_push
        if not host:
            self._after_torch()
        self._ck(fn(self.h, p, ld, kind))
        # the i-th push after data_init serves step i as the current snapshot and step i+1 as the previous one
        self._npush += 1
        self._borrow(keep, self._npush + 1)

    # -- DE problems -------------------------------------------------------------------------------
    def set_substepper(self, flow, ode, nsub=1, abstol=0.0, reltol=0.0, maxiters=None):
        self._ck(self.lib.dlra_set_substepper(self.h, flow, ode, nsub, abstol, reltol))
        if maxiters is not None:
            self._ck(self.lib.dlra_set_substepper_maxiters(self.h, flow, int(maxiters)))

    @staticmethod
    def _operator(x, keep):
        """dlra_operator for a Python scalar (s*I), a CSR tuple (rowptr int64, colind int32, values fp64, shape) of CUDA tensors,
        or a dense CUDA matrix; tensors that must stay alive are appended to `keep`."""
        if x is None:
            return None
        o = L.Operator()
        if np.isscalar(x):
            o.kind, o.scale = L.OP_IDENTITY_SCALED, float(x)
            return o
        if isinstance(x, tuple):
            rowptr, colind, values, shape = x
            keep.extend([rowptr, colind, values])
            o.kind, o.rows, o.cols = L.OP_CSR, shape[0], shape[1]
            o.rowptr, o.colind, o.values, o.scale = rowptr.data_ptr(), colind.data_ptr(), values.data_ptr(), 1.0
            return o
        x = colmajor_device(x)
        keep.append(x)
        o.kind, o.rows, o.cols, o.dense, o.scale = L.OP_DENSE, x.shape[0], x.shape[1], x.data_ptr(), 1.0
        o.ld = x.stride(1) if x.shape[1] > 1 else x.shape[0]
        return o

    def rhs_set(self, A=None, B=None, G=None, H=None, D1=None, D2=None, c_had=0.0):
        keep = []
        ops = [self._operator(x, keep) for x in (A, B, D1, D2)]
        refs = [C.byref(o) if o is not None else None for o in ops]
        q = 0
        pg = ph = None
        ldg = ldh = 0
        if G is not None:
            G, H = colmajor_device(G), colmajor_device(H)
            keep.extend([G, H])
            q = G.shape[1]
            pg, ldg = G.data_ptr(), (G.stride(1) if q > 1 else G.shape[0])
            ph, ldh = H.data_ptr(), (H.stride(1) if q > 1 else H.shape[0])
        self._after_torch()  # after every layout conversion above has been queued on torch's stream
        self._ck(self.lib.dlra_rhs_set(self.h, refs[0], refs[1], pg, ldg, ph, ldh, q, refs[2], refs[3], float(c_had)))
        self._keep["rhs"] = (keep, ops)
        self._keep["rhs_terms"] = []

    def rhs_add_term(self, A, B):
        """F(X) += A*X*B' (two-sided term; A, B as accepted by rhs_set, a scalar s meaning s*I)."""
        keep = []
        oa, ob = self._operator(A, keep), self._operator(B, keep)
        assert oa is not None and ob is not None, "a two-sided term needs both operators (use 1.0 for the identity)"
        self._after_torch()
        self._ck(self.lib.dlra_rhs_add_term(self.h, C.byref(oa), C.byref(ob)))
        self._keep.setdefault("rhs_terms", []).append((keep, oa, ob))

    # -- steps -------------------------------------------------------------------------------------
    def step_ksl(self, order, t=0.0, dt=1.0):
        self._ck(self.lib.dlra_step_ksl(self.h, order, float(t), float(dt)))

    def step_bug(self, t=0.0, dt=1.0):
        self._ck(self.lib.dlra_step_bug(self.h, float(t), float(dt)))

    def step_rabug(self, tol, rmax, t=0.0, dt=1.0):
        rn, ch = C.c_int(), C.c_int()
        self._ck(self.lib.dlra_step_rabug(self.h, float(t), float(dt), float(tol), int(min(rmax, 2 ** 62)), C.byref(rn), C.byref(ch)))
        return rn.value, bool(ch.value)

    def step_greedy(self, t=0.0, dt=1.0):
        self._ck(self.lib.dlra_step_greedy(self.h, float(t), float(dt)))

    def step_greedy_two_factor(self, mode, t=0.0, dt=1.0, carry_fsal=True):
        self._ck(self.lib.dlra_step_greedy_two_factor(self.h, int(mode), 1 if carry_fsal else 0, float(t), float(dt)))

    def sync(self):
        self._ck(self.lib.dlra_sync(self.h))

    # -- diagnostics -------------------------------------------------------------------------------
    def reconstruct_error(self, Yref):
        p, ld, host, keep = _ptr_ld(Yref)
        assert not host, "reconstruct_error takes a device matrix"
        self._after_torch()
        out = C.c_double()
        self._ck(self.lib.dlra_reconstruct_error(self.h, p, ld, C.byref(out)))
        return out.value

    def normal_component(self, dY, C_mat=None, tol=1e-8, want_matrix=False):
        """‖N‖_F (and N as a device matrix when want_matrix) of N = (I-UU') dY (I - Z pinv(C, atol=tol) Z') (utils.jl:2-20)."""
        p, ld, host, keep = _ptr_ld(dY)
        assert not host, "normal_component takes a device matrix"
        pc, ldc = None, 0
        if C_mat is not None:
            C_mat = colmajor_device(C_mat)
            pc, ldc = C_mat.data_ptr(), (C_mat.stride(1) if C_mat.shape[1] > 1 else C_mat.shape[0])
        out = empty_colmajor(self.n, self.m, torch.device("cuda", self.device)) if want_matrix else None
        self._after_torch()
        nrm = C.c_double()
        self._ck(self.lib.dlra_normal_component(self.h, p, ld, pc, ldc, float(tol), out.data_ptr() if want_matrix else None, self.n,
                                                C.byref(nrm)))
        return (nrm.value, out) if want_matrix else nrm.value

    def reconstruct(self):
        Y = empty_colmajor(self.n, self.m, torch.device("cuda", self.device))
        self._ck(self.lib.dlra_reconstruct(self.h, Y.data_ptr(), self.n))
        self.sync()
        return Y

    def set_profiling(self, on=True):
        self._ck(self.lib.dlra_set_profiling(self.h, 1 if on else 0))

    def stats(self, reset=False):
        kl, pl = C.c_int64(), C.c_int64()
        ms, by = C.c_double(), C.c_double()
        self._ck(self.lib.dlra_stats(self.h, C.byref(kl), C.byref(pl), C.byref(ms), C.byref(by), 1 if reset else 0))
        return {"kernel_launches": kl.value, "pass_launches": pl.value, "pass_ms": ms.value, "pass_bytes": by.value}

    def pass_breakdown(self):
        la = (C.c_int64 * 4)()
        ms, by, fl = (C.c_double * 4)(), (C.c_double * 4)(), (C.c_double * 4)()
        self._ck(self.lib.dlra_pass_breakdown(self.h, la, ms, by, fl))
        names = ("fused_KL", "K_only", "L_only", "pipelined_SKL")
        return {nm: {"launches": la[i], "ms": ms[i], "bytes": by[i], "flops": fl[i]} for i, nm in enumerate(names)}

    def event_record(self, slot):
        self._ck(self.lib.dlra_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        out = C.c_double()
        self._ck(self.lib.dlra_event_elapsed_ms(self.h, a, b, C.byref(out)))
        return out.value
