"""Engine: one `pvw_ctx` (one CUDA device, one shard of parties) behind an object, speaking numpy on the host
side and torch CUDA tensors / raw device pointers on the device side.  No arithmetic happens here: every method
is one call into the C ABI (include/pvw_b200.h)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from .errors import PvwError


def _is_device_tensor(x) -> bool:
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


class _Arg:
    """pointer + keep-alive for one call argument (numpy host array or torch CUDA tensor)"""

    def __init__(self, x, dtype, shape: Optional[Sequence[int]] = None, name: str = "argument"):
        self.device = _is_device_tensor(x)
        if self.device:
            import torch
            want = {np.uint64: (torch.uint64, torch.int64), np.int64: (torch.int64,), np.uint32: (torch.uint32, torch.int32)}[dtype]
            if x.dtype not in want or not x.is_contiguous():
                raise PvwError("InvalidParameters", f"{name}: need a contiguous CUDA tensor of dtype {want[0]}")
            self.keep = x
            self.ptr = x.data_ptr()
            got = tuple(x.shape)
        else:
            a = np.ascontiguousarray(x, dtype=dtype)
            self.keep = a
            self.ptr = a.ctypes.data
            got = a.shape
        if shape is not None and tuple(got) != tuple(shape):
            raise PvwError("DimensionMismatch", f"{name}: expected shape {tuple(shape)}, got {tuple(got)}")


class _SmallArg:
    """Small signed integers (secrets, randomness, errors): int64 -- the reference's i64 -- or the narrow forms the C ABI accepts
    (PVW_IN_SECRET_I8 for r / sk, PVW_IN_ERROR_I32 / _I16 for e, e1, e2).  The element type of the array picks the flag."""

    def __init__(self, x, shape, name: str, kind: str):
        allowed = {"secret": (1, 8), "error": (2, 4, 8)}[kind]
        self.device = _is_device_tensor(x)
        if self.device:
            import torch
            sizes = {torch.int8: 1, torch.int16: 2, torch.int32: 4, torch.int64: 8}
            if x.dtype not in sizes or not x.is_contiguous():
                raise PvwError("InvalidParameters", f"{name}: need a contiguous signed-integer CUDA tensor")
            self.bytes = sizes[x.dtype]
            self.keep, self.ptr, got = x, x.data_ptr(), tuple(x.shape)
        else:
            a = np.asarray(x)
            if a.dtype not in (np.int8, np.int16, np.int32, np.int64) or a.dtype.itemsize not in allowed:
                a = a.astype(np.int64)
            a = np.ascontiguousarray(a)
            self.bytes = a.dtype.itemsize
            self.keep, self.ptr, got = a, a.ctypes.data, a.shape
        if self.bytes not in allowed:
            raise PvwError("InvalidParameters", f"{name}: element size {self.bytes} is not accepted for {kind} inputs (allowed: {allowed} bytes)")
        if tuple(got) != tuple(shape):
            raise PvwError("DimensionMismatch", f"{name}: expected shape {tuple(shape)}, got {tuple(got)}")
        self.flag = {("secret", 1): _ffi.PVW_IN_SECRET_I8, ("error", 4): _ffi.PVW_IN_ERROR_I32, ("error", 2): _ffi.PVW_IN_ERROR_I16}.get((kind, self.bytes), 0)


def _same_error_type(*args):
    sizes = {a.bytes for a in args if a is not None}
    if len(sizes) > 1:
        raise PvwError("InvalidParameters", "error inputs of one call must share one element type")


class Engine:
    """Owns a pvw_ctx.  `row0`/`nrows` select the shard of parties (rows of B) this context holds."""

    def __init__(self, n: int, k: int, l: int, moduli: Sequence[int], psi: Optional[Sequence[int]] = None,
                 secret_variance: float = 0.5, error_bound_1: int = 100, error_bound_2: int = 200,
                 row0: int = 0, nrows: int = 0, device: int = 0):
        self.lib = _ffi.load()
        for name, v in (("n", n), ("k", k), ("l", l), ("row0", row0), ("nrows", nrows)):
            if not 0 <= int(v) < 2 ** 32:
                raise PvwError("InvalidParameters", f"{name} out of range")
        for name, v in (("error_bound_1", error_bound_1), ("error_bound_2", error_bound_2)):
            if not 0 <= int(v) < 2 ** 63:
                # the reference type is BigInt (parameters.rs:31-33); every call site uses <= u32 (parameters.rs:110-114)
                raise PvwError("InvalidParameters", f"{name} must be in [0, 2^63)")
        mods = [int(q) for q in moduli]
        if any(q < 0 or q >= 2 ** 64 for q in mods):
            raise PvwError("InvalidParameters", "Context creation failed: modulus does not fit 64 bits")
        if psi is not None and len(psi) != len(mods):
            raise PvwError("InvalidParameters", f"{len(psi)} roots given for {len(mods)} moduli")
        self._mods = (C.c_uint64 * max(1, len(mods)))(*mods)
        self._psi = (C.c_uint64 * len(mods))(*[int(p) for p in psi]) if psi is not None else None
        desc = _ffi.PvwParamsDesc(int(n), int(k), int(l), len(mods), self._mods if mods else None, self._psi,
                                  float(secret_variance), int(error_bound_1), int(error_bound_2), int(row0), int(nrows), int(device))
        h = C.c_void_p()
        rc = self.lib.pvw_ctx_create(C.byref(h), C.byref(desc))
        if rc != 0:
            raise PvwError(_ffi.STATUS_NAMES.get(rc, "InternalError"), (self.lib.pvw_last_error(None) or b"").decode())
        self.h = h
        self.n, self.k, self.l, self.L = int(n), int(k), int(l), len(mods)
        self.moduli = mods
        self.row0 = int(row0)
        self.nrows = int(nrows) if nrows else self.n - self.row0
        self.device = int(device)
        self.poly = (self.L, self.l)
        self.capacity = 0

    # -- plumbing ---------------------------------------------------------------------------------
    def _torch_streams(self):
        import torch
        if getattr(self, "_ext", None) is None:
            self._ext = torch.cuda.ExternalStream(self.stream, device=torch.device(f"cuda:{self.device}"))
        return self._ext, torch.cuda.current_stream(torch.device(f"cuda:{self.device}"))

    def _before_device_call(self):
        """CUDA tensors handed to the library were produced on torch's current stream; the library works on its own
        stream (pvw_ctx_stream).  Order the library's stream after torch's so that it never reads unfinished inputs."""
        ext, cur = self._torch_streams()
        if ext.cuda_stream != cur.cuda_stream:
            ext.wait_stream(cur)

    def _after_device_call(self):
        """... and torch's current stream after the library's, so that torch ops see finished outputs and do not recycle
        input buffers the library is still reading."""
        ext, cur = self._torch_streams()
        if ext.cuda_stream != cur.cuda_stream:
            cur.wait_stream(ext)

    def _check(self, rc: int):
        if rc != 0:
            raise PvwError(_ffi.STATUS_NAMES.get(rc, "InternalError"), (self.lib.pvw_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.pvw_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._check(self.lib.pvw_ctx_synchronize(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.pvw_ctx_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.pvw_ctx_launch_count(self.h))

    def set_option(self, name: str, value: int):
        self._check(self.lib.pvw_ctx_set_option(self.h, name.encode(), int(value)))

    def profile(self) -> dict:
        """{kernel kind: (ms_total, launches, algorithmic_bytes)} since profiling was last reset (option "profile")"""
        out = {}
        for i, name in enumerate(_ffi.KERNEL_KINDS):
            ms, n, b = C.c_double(), C.c_uint64(), C.c_double()
            self._check(self.lib.pvw_ctx_profile(self.h, i, C.byref(ms), C.byref(n), C.byref(b)))
            out[name] = (ms.value, n.value, b.value)
        return out

    # -- parameters ------------------------------------------------------------------------------
    def _bigint(self, which: int) -> int:
        nw = C.c_uint32()
        self._check(self.lib.pvw_params_bigint(self.h, which, None, 0, C.byref(nw)))
        buf = np.zeros(max(1, nw.value), dtype=np.uint64)
        self._check(self.lib.pvw_params_bigint(self.h, which, buf.ctypes.data, len(buf), C.byref(nw)))
        return sum(int(w) << (64 * i) for i, w in enumerate(buf[:nw.value]))

    @property
    def q_total(self) -> int:
        return self._bigint(0)

    @property
    def delta(self) -> int:
        return self._bigint(1)

    @property
    def delta_power_l_minus_1(self) -> int:
        return self._bigint(2)

    @property
    def psi(self):
        out = np.zeros(self.L, dtype=np.uint64)
        self._check(self.lib.pvw_params_psi(self.h, out.ctypes.data))
        return [int(x) for x in out]

    def verify_correctness_condition(self) -> bool:
        ok = C.c_int()
        self._check(self.lib.pvw_params_correctness_condition(self.h, C.byref(ok)))
        return bool(ok.value)

    # -- CRS / public key residency --------------------------------------------------------------
    def crs_upload(self, A):
        a = _Arg(A, np.uint64, (self.k, self.k) + self.poly, "A")
        if a.device:
            self._before_device_call()
        self._check(self.lib.pvw_crs_upload(self.h, a.ptr, _ffi.PVW_IO_DEVICE if a.device else 0))
        if a.device:
            self._after_device_call()

    def crs_generate_deterministic(self, seed: bytes, want_matrix: bool = False):
        """PvwCrs::new_deterministic (crs.rs:45-67): expand the 32-byte master seed on the host and upload"""
        seed = bytes(seed)
        if len(seed) != 32:
            raise PvwError("InvalidParameters", "the master seed must be 32 bytes")
        out = np.empty((self.k, self.k) + self.poly, dtype=np.uint64) if want_matrix else None
        self._check(self.lib.pvw_crs_generate_deterministic(self.h, seed, out.ctypes.data if want_matrix else None))
        return out

    def crs_generate_from_tag(self, tag: str, want_matrix: bool = False):
        """PvwCrs::new_from_tag (crs.rs:74-90)"""
        out = np.empty((self.k, self.k) + self.poly, dtype=np.uint64) if want_matrix else None
        self._check(self.lib.pvw_crs_generate_from_tag(self.h, tag.encode(), out.ctypes.data if want_matrix else None))
        return out

    def crs_download(self) -> np.ndarray:
        out = np.empty((self.k, self.k) + self.poly, dtype=np.uint64)
        self._check(self.lib.pvw_crs_download(self.h, out.ctypes.data))
        return out

    def pk_upload_rows(self, row: int, B):
        count = int(B.shape[0])
        a = _Arg(B, np.uint64, (count, self.k) + self.poly, "B rows")
        if a.device:
            self._before_device_call()
        self._check(self.lib.pvw_pk_upload_rows(self.h, row, count, a.ptr, _ffi.PVW_IO_DEVICE if a.device else 0))
        if a.device:
            self._after_device_call()

    def pk_download_rows(self, row: int, count: int) -> np.ndarray:
        out = np.empty((count, self.k) + self.poly, dtype=np.uint64)
        self._check(self.lib.pvw_pk_download_rows(self.h, row, count, out.ctypes.data))
        return out

    @property
    def num_keys(self) -> int:
        v = C.c_uint32()
        self._check(self.lib.pvw_pk_num_keys(self.h, C.byref(v)))
        return v.value

    def keygen_batch(self, row: int, sk, e):
        count = int(sk.shape[0])
        a = _SmallArg(sk, (count, self.k, self.l), "sk", "secret")
        b = _SmallArg(e, (count, self.k, self.l), "e", "error")
        if a.device != b.device:
            raise PvwError("InvalidParameters", "sk and e must both be host or both be device arrays")
        if a.device:
            self._before_device_call()
        self._check(self.lib.pvw_keygen_batch(self.h, row, count, a.ptr, b.ptr, (_ffi.PVW_IO_DEVICE if a.device else 0) | a.flag | b.flag))
        if a.device:
            self._after_device_call()

    def crs_multiply_by_randomness(self, r_hat) -> np.ndarray:
        r = np.ascontiguousarray(r_hat, dtype=np.uint64)
        D = r.shape[0]
        if r.shape != (D, self.k) + self.poly:
            raise PvwError("DimensionMismatch", f"expected {self.k} polynomials, got shape {r.shape}")
        out = np.empty_like(r)
        self._check(self.lib.pvw_crs_multiply_by_randomness(self.h, D, r.ctypes.data, out.ctypes.data))
        return out

    # -- ciphertext store ------------------------------------------------------------------------
    def ct_reserve(self, capacity: int):
        self._check(self.lib.pvw_ct_reserve(self.h, capacity))
        self.capacity = capacity

    def encrypt_batch(self, slot0: int, m, r, e1, e2, c1_range=None, part: str = "both", push_c1: bool = False):
        """m [D][nrows] u64; r, e1 [D][k][l] i64; e2 [D][nrows][l] i64 -- all host or all device.
        part: "both", or "c1" / "c2" alone (PVW_ENC_C1_ONLY / PVW_ENC_C2_ONLY: a multi-GPU host layer all-gathers the c1 slices
        while the c2 product runs); m / e2 may be None with "c1", e1 with "c2".
        push_c1: on a connected shard exchange, queue the peer copies of the c1 slice inside the call (PVW_ENC_PUSH_C1)."""
        D = int(r.shape[0])
        lo, hi = (0, D) if c1_range is None else c1_range
        pflag = {"both": 0, "c1": _ffi.PVW_ENC_C1_ONLY, "c2": _ffi.PVW_ENC_C2_ONLY}[part] | (_ffi.PVW_ENC_PUSH_C1 if push_c1 else 0)
        am = _Arg(m, np.uint64, (D, self.nrows), "m") if m is not None else None
        ar = _SmallArg(r, (D, self.k, self.l), "r", "secret")
        ae2 = _SmallArg(e2, (D, self.nrows, self.l), "e2", "error") if e2 is not None else None
        ae1 = _SmallArg(e1, (D, self.k, self.l), "e1", "error") if e1 is not None else None
        _same_error_type(ae1, ae2)
        if part != "c1" and (am is None or ae2 is None):
            raise PvwError("InvalidParameters", "m and e2 are required")
        devs = {a.device for a in (am, ar, ae2, ae1) if a is not None}
        if len(devs) != 1:
            raise PvwError("InvalidParameters", "inputs must be all host or all device arrays")
        on_device = devs.pop()
        if on_device:
            self._before_device_call()
        self._check(self.lib.pvw_encrypt_batch(self.h, slot0, D, lo, hi, am.ptr if am else None, ar.ptr, ae1.ptr if ae1 else None,
                                               ae2.ptr if ae2 else None,
                                               (_ffi.PVW_IO_DEVICE if on_device else 0) | pflag | ar.flag | (ae1.flag if ae1 else ae2.flag if ae2 else 0)))
        if on_device:
            self._after_device_call()

    def ct_download(self, slot: int, want_c1=True, want_c2=True):
        c1 = np.empty((self.k,) + self.poly, dtype=np.uint64) if want_c1 else None
        c2 = np.empty((self.nrows,) + self.poly, dtype=np.uint64) if want_c2 else None
        self._check(self.lib.pvw_ct_download(self.h, slot, c1.ctypes.data if want_c1 else None, c2.ctypes.data if want_c2 else None))
        return c1, c2

    def ct_upload(self, slot: int, c1=None, c2=None):
        a1 = _Arg(c1, np.uint64, (self.k,) + self.poly, "c1") if c1 is not None else None
        a2 = _Arg(c2, np.uint64, (self.nrows,) + self.poly, "c2") if c2 is not None else None
        self._check(self.lib.pvw_ct_upload(self.h, slot, a1.ptr if a1 else None, a2.ptr if a2 else None))

    def c1_device_ptr(self, slot: int = 0):
        p, stride = C.c_void_p(), C.c_uint64()
        self._check(self.lib.pvw_ct_c1_device_ptr(self.h, slot, C.byref(p), C.byref(stride)))
        return int(p.value), int(stride.value)

    def c1_store_tensor(self, slot0: int, count: int):
        """torch view [count][L*k*l] (int64 reinterpretation) of the device-resident c1 of slots [slot0, slot0+count),
        for collectives issued by the host layer (NCCL all-gather over dealers)."""
        import torch
        ptr, stride = self.c1_device_ptr(slot0)

        class _Holder:
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (count, stride), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}
        return torch.as_tensor(h, device=f"cuda:{self.device}")

    # -- multi-GPU c1 exchange over the copy engines (pvw_shard_*; host logic in sharding.CopyEngineExchange) ----
    def shard_export(self, world: int) -> bytes:
        h = _ffi.PvwShardHandle()
        self._check(self.lib.pvw_shard_export(self.h, int(world), C.byref(h)))
        return bytes(h.bytes)

    def shard_connect(self, world: int, rank: int, handles: Sequence[bytes]):
        if len(handles) != world or any(len(b) != 192 for b in handles):
            raise PvwError("InvalidParameters", f"need {world} handles of 192 bytes")
        arr = (_ffi.PvwShardHandle * world)()
        for i, b in enumerate(handles):
            C.memmove(C.byref(arr[i]), bytes(b), 192)
        self._check(self.lib.pvw_shard_connect(self.h, int(world), int(rank), arr))

    def shard_push_c1(self, slot0: int, count: int):
        self._check(self.lib.pvw_shard_push_c1(self.h, int(slot0), int(count)))

    def shard_wait_c1(self):
        self._check(self.lib.pvw_shard_wait_c1(self.h))

    def shard_release_c1(self):
        self._check(self.lib.pvw_shard_release_c1(self.h))

    def shard_disconnect(self):
        self._check(self.lib.pvw_shard_disconnect(self.h))

    # -- wire format: bincode of the crate's serde impls, produced / parsed on the device ---------
    @property
    def wire_layout(self) -> "_ffi.PvwWireLayout":
        if getattr(self, "_wire", None) is None:
            w = _ffi.PvwWireLayout()
            self._check(self.lib.pvw_wire_layout_get(self.h, C.byref(w)))
            self._wire = w
        return self._wire

    def wire_params(self) -> bytes:
        """bincode::serialize(&PvwParameters)  (parameters.rs:606-623)"""
        buf = np.empty(self.wire_layout.params_bytes, dtype=np.uint8)
        self._check(self.lib.pvw_wire_params(self.h, buf.ctypes.data, buf.size))
        return buf.tobytes()

    def _bytes_arg(self, x, nbytes: int, name: str):
        """(pointer, keep-alive, on_device) of a byte buffer: numpy / bytes on the host or a torch.uint8 CUDA tensor"""
        if _is_device_tensor(x):
            import torch
            if x.dtype != torch.uint8 or not x.is_contiguous():
                raise PvwError("InvalidParameters", f"{name}: need a contiguous CUDA tensor of dtype uint8")
            if x.numel() < nbytes:
                raise PvwError("InsufficientData", f"{name}: expected {nbytes} bytes, got {x.numel()}")
            return x.data_ptr(), x, True
        a = np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray, memoryview)) else np.ascontiguousarray(x, dtype=np.uint8)
        if a.size < nbytes:
            raise PvwError("InsufficientData", f"{name}: expected {nbytes} bytes, got {a.size}")
        return a.ctypes.data, a, False

    def wire_ct_serialize(self, slot0: int, D: int, out=None, stride: Optional[int] = None):
        """D stored ciphertexts -> D bincode(PvwCiphertext) blobs (encryption.rs:298-317), `stride` bytes apart.
        out: None (a new host uint8 array [D][stride] is returned) or a torch.uint8 CUDA tensor (stays in HBM)."""
        stride = int(stride or self.wire_layout.ciphertext_bytes)
        if out is None:
            out = np.empty((D, stride), dtype=np.uint8)
        ptr, keep, dev = self._bytes_arg(out, D * stride, "out")
        if dev:
            self._before_device_call()
        self._check(self.lib.pvw_wire_ct_serialize(self.h, slot0, D, ptr, stride, _ffi.PVW_IO_DEVICE if dev else 0))
        if dev:
            self._after_device_call()
        return out

    def wire_ct_deserialize(self, slot0: int, D: int, blobs, stride: Optional[int] = None):
        stride = int(stride or self.wire_layout.ciphertext_bytes)
        ptr, keep, dev = self._bytes_arg(blobs, D * stride if D else 0, "blobs")
        if dev:
            self._before_device_call()
        self._check(self.lib.pvw_wire_ct_deserialize(self.h, slot0, D, ptr, stride, _ffi.PVW_IO_DEVICE if dev else 0))
        if dev:
            self._after_device_call()

    def wire_pk_serialize_rows(self, row: int, count: int, out=None):
        """rows of GlobalPublicKey.matrix as `count` consecutive Vec<Vec<u8>> (public_key.rs:528-533)"""
        nb = count * self.wire_layout.pk_row_bytes
        if out is None:
            out = np.empty(nb, dtype=np.uint8)
        ptr, keep, dev = self._bytes_arg(out, nb, "out")
        if dev:
            self._before_device_call()
        self._check(self.lib.pvw_wire_pk_serialize_rows(self.h, row, count, ptr, _ffi.PVW_IO_DEVICE if dev else 0))
        if dev:
            self._after_device_call()
        return out

    def wire_pk_deserialize_rows(self, row: int, count: int, data):
        ptr, keep, dev = self._bytes_arg(data, count * self.wire_layout.pk_row_bytes, "data")
        if dev:
            self._before_device_call()
        self._check(self.lib.pvw_wire_pk_deserialize_rows(self.h, row, count, ptr, _ffi.PVW_IO_DEVICE if dev else 0))
        if dev:
            self._after_device_call()

    def wire_crs_serialize(self) -> bytes:
        """bincode::serialize(&PvwCrs)  (crs.rs:228-249)"""
        buf = np.empty(self.wire_layout.crs_bytes, dtype=np.uint8)
        self._check(self.lib.pvw_wire_crs_serialize(self.h, buf.ctypes.data, buf.size, 0))
        return buf.tobytes()

    def wire_crs_deserialize(self, data):
        a = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, dtype=np.uint8)
        self._check(self.lib.pvw_wire_crs_deserialize(self.h, a.ctypes.data, a.size, 0))

    def wire_polys_serialize(self, polys) -> bytes:
        """[count][L][l] residues -> count records (u64 length + Poly::to_bytes each)"""
        a = _Arg(polys, np.uint64, None, "polys")
        if a.device or a.keep.ndim != 3 or a.keep.shape[1:] != self.poly:
            raise PvwError("DimensionMismatch", f"polys: expected a host array [count]{list(self.poly)}")
        count = a.keep.shape[0]
        buf = np.empty(count * self.wire_layout.record_bytes, dtype=np.uint8)
        self._check(self.lib.pvw_wire_polys_serialize(self.h, count, a.ptr, buf.ctypes.data))
        return buf.tobytes()

    def wire_polys_deserialize(self, data, count: int) -> np.ndarray:
        ptr, keep, dev = self._bytes_arg(data, count * self.wire_layout.record_bytes, "data")
        if dev:
            raise PvwError("InvalidParameters", "data: host bytes expected")
        out = np.empty((count,) + self.poly, dtype=np.uint64)
        self._check(self.lib.pvw_wire_polys_deserialize(self.h, count, ptr, out.ctypes.data))
        return out

    def decrypt_batch(self, party_idx, sk, dealer_slots=None, D: Optional[int] = None, out=None):
        """out[p][d] for P parties (global indices inside the shard) x D stored ciphertexts"""
        pidx = np.ascontiguousarray(party_idx, dtype=np.uint32)
        P = len(pidx)
        if dealer_slots is not None:
            ds = np.ascontiguousarray(dealer_slots, dtype=np.uint32)
            D = len(ds)
        else:
            ds = None
            D = self.capacity if D is None else D
        a = _SmallArg(sk, (P, self.k, self.l), "sk", "secret")
        if a.device:
            if out is None:
                import torch
                out = torch.empty((P, D), dtype=torch.int64, device=sk.device)
            o = _Arg(out, np.uint64, (P, D), "out")
            self._before_device_call()
            self._check(self.lib.pvw_decrypt_batch(self.h, D, ds.ctypes.data if ds is not None else None, P, pidx.ctypes.data, a.ptr, o.ptr,
                                                   _ffi.PVW_IO_DEVICE | a.flag))
            self._after_device_call()
            return out
        if out is not None:      # caller-provided host buffer (e.g. pinned memory): no allocation / page faults per call
            res = out
            if not (isinstance(res, np.ndarray) and res.dtype == np.uint64 and res.shape == (P, D) and res.flags.c_contiguous):
                raise PvwError("DimensionMismatch", f"out: need a C-contiguous uint64 array of shape {(P, D)}")
        else:
            res = np.empty((P, D), dtype=np.uint64)
        self._check(self.lib.pvw_decrypt_batch(self.h, D, ds.ctypes.data if ds is not None else None, P, pidx.ctypes.data, a.ptr,
                                               res.ctypes.data, a.flag))
        return res

    def decode_batch(self, zhat) -> np.ndarray:
        z = np.ascontiguousarray(zhat, dtype=np.uint64)
        count = z.size // (self.L * self.l)
        out = np.empty(count, dtype=np.uint64)
        self._check(self.lib.pvw_decode_batch(self.h, count, z.ctypes.data, out.ctypes.data))
        return out

    def encode_scalars(self, scalars) -> np.ndarray:
        m = np.ascontiguousarray(scalars, dtype=np.uint64)
        out = np.empty(m.shape + self.poly, dtype=np.uint64)
        self._check(self.lib.pvw_encode_scalars(self.h, m.size, m.ctypes.data, out.ctypes.data))
        return out

    def ntt_forward_small(self, coeffs) -> np.ndarray:
        c = np.ascontiguousarray(coeffs, dtype=np.int64)
        if c.shape[-1] != self.l:
            raise PvwError("InvalidParameters", f"Expected {self.l} coefficients, got {c.shape[-1]}")
        lead = c.shape[:-1]
        out = np.empty(lead + self.poly, dtype=np.uint64)
        self._check(self.lib.pvw_ntt_forward_small(self.h, int(np.prod(lead, dtype=np.int64)), c.ctypes.data, out.ctypes.data))
        return out
